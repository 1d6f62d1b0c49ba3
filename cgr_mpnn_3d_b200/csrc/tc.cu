// tcgen05 engine (CGR_ENGINE_TC): the forward of cgr_mpnn_3D/models/GNN.py:76-110 on 5th-gen tensor cores.
//
// Formulation.  Every Linear of the path is linear, so the gather commutes with it:
//     W (sum_{k in in(src e)} h[k] - h[e^1])  =  sum_{k in in(src e)} (W h[k]) - W h[e^1].
// Each layer is therefore ONE kernel: a TMA-fed tcgen05 GEMM  y = h W^T  of a 128-bond row tile with the
// accumulator in TMEM, whose epilogue stages y in shared memory and performs the directed-bond gather
// there (index rows staged in shared memory too), then adds bias + skip*h0, applies the activation /
// dropout and writes the next layer's operand.  A row tile holds WHOLE reactions (tile plan), so every
// gather is tile-local.
//
// Precision.  fp32 parity (1e-4 on Ea) is kept with an FP16x3 split: v = hi + lo (two fp16 numbers,
// 22 significant bits), a.b ~= ahi.bhi + alo.bhi + ahi.blo accumulated in fp32 by three tcgen05.mma per
// k-step.  Weights are pre-scaled by a power of two per matrix so their lo parts stay in the fp16
// normal range; the scale is undone (exactly) in the epilogue.
#include "tc.cuh"

#include <math.h>
#include <mutex>
#include <unordered_map>
#include <stdlib.h>
#include <string.h>

#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"
#include "tc_fwd.cuh"
#include "tc_proj.cuh"
#include "umma.cuh"

namespace {

using namespace tcg;

// ------------------------------------------------------------------------------------------------
// small SIMT kernels around the GEMMs
// ------------------------------------------------------------------------------------------------

// fp32 rows -> (hi, lo) fp16 rows; one warp per row
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ in, int64_t ld, int64_t rows, int cols,
                                                         __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldo,
                                                         int* __restrict__ overflow, int flag_bit) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  bool ovf = false;
  for (int c = threadIdx.x & 31; c < cols; c += 32) {
    const float v = __ldg(in + r * ld + c);
    ovf |= fabsf(v) > 60000.f;
    __half h, l;
    split_f16(v, h, l);
    hi[r * ldo + c] = h;
    lo[r * ldo + c] = l;
  }
  if (ovf && overflow) atomicOr(overflow, flag_bit);
}

// tc_status[0] bits (cgr_b200.h): bit 0 = an activation of THIS forward left the fp16 range (cleared when a forward
// starts), bit 1 = a feature of data.x did (set by the feature split at batch preparation, kept for the batch's life)
constexpr int FLAG_ACT = 1, FLAG_X = 2;
__global__ void flag_begin_kernel(int* flag) { atomicAnd(flag, ~FLAG_ACT); }
// energies computed from clipped operands are wrong: make that impossible to miss (NaN), no host synchronisation needed
__global__ void poison_kernel(float* out, int64_t n, const int* flag) {
  if ((*flag & (FLAG_ACT | FLAG_X)) == 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __int_as_float(0x7fc00000);
}

// h0 = act(P'[src e] + ea[e] . W_e^T) on tile-packed rows (P' already holds b_i).  A block owns 32 bond rows of one
// tile: W_e^T [fb][H] and the rows' bond features are staged in shared memory, each warp then produces 4 rows at a
// time (4 independent gathers of P' in flight per lane).  Writes the fp32 copy (skip operand of every layer) and the
// FP16 (hi, lo) operand of layer 0.
constexpr int EI_ROWS = 32;
__global__ void __launch_bounds__(256) tc_edge_init_kernel(const float* __restrict__ PQ, int64_t ldpq,
                                                           const float* __restrict__ ea, const int32_t* __restrict__ src,
                                                           const float* __restrict__ wet,
                                                           const int32_t* __restrict__ tile_info, int fb, int H, int act,
                                                           float* __restrict__ h0, __half* __restrict__ o_hi,
                                                           __half* __restrict__ o_lo, int64_t ldo,
                                                           int* __restrict__ overflow) {
  extern __shared__ __align__(16) float ei_smem[];
  float* wet_s = ei_smem;                          // [fb][H]
  float* ea_s = ei_smem + (size_t)fb * H;          // [EI_ROWS][fb]
  const int tile = blockIdx.y;
  const int ebase = __ldg(tile_info + tile * 8), ecount = __ldg(tile_info + tile * 8 + 1);
  const int j0 = blockIdx.x * EI_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  umma::grid_dep_launch();
  if (j0 >= ecount) return;
  // weights and bond features do not depend on the previous kernel: stage them before the dependency wait
  for (int i = threadIdx.x; i < fb * H / 4; i += blockDim.x)
    reinterpret_cast<float4*>(wet_s)[i] = __ldg(reinterpret_cast<const float4*>(wet) + i);
  for (int i = threadIdx.x; i < EI_ROWS * fb; i += blockDim.x) {
    const int j = j0 + i / fb;
    ea_s[i] = j < ecount ? __ldg(ea + (int64_t)(ebase + j) * fb + (i % fb)) : 0.f;
  }
  umma::grid_dep_wait();                          // PQ comes from the atom-projection kernel
  __syncthreads();
  const int jw = j0 + warp * 4;                    // this warp's 4 rows
  if (jw >= ecount) return;
  const float* prow[4];
  int64_t r[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = jw + u < ecount ? jw + u : jw;
    prow[u] = PQ + (int64_t)__ldg(src + ebase + j) * ldpq;
    r[u] = (int64_t)tile * TM + j;
  }
  float vmax = 0.f;
  for (int n = 4 * lane; n < H; n += 128) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = __ldg(reinterpret_cast<const float4*>(prow[u] + n));
    for (int k = 0; k < fb; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(wet_s + (size_t)k * H + n);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float ek = ea_s[(warp * 4 + u) * fb + k];
        a[u].x = fmaf(ek, w.x, a[u].x); a[u].y = fmaf(ek, w.y, a[u].y);
        a[u].z = fmaf(ek, w.z, a[u].z); a[u].w = fmaf(ek, w.w, a[u].w);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (jw + u >= ecount) break;
      float4 v = a[u];
      v.x = cgr_act(v.x, act); v.y = cgr_act(v.y, act); v.z = cgr_act(v.z, act); v.w = cgr_act(v.w, act);
      vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      *reinterpret_cast<float4*>(h0 + r[u] * H + n) = v;
      const __half2 hi01 = __floats2half2_rn(v.x, v.y), hi23 = __floats2half2_rn(v.z, v.w);
      const float2 f01 = __half22float2(hi01), f23 = __half22float2(hi23);
      const __half2 lo01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), lo23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
      uint2 ph, pl;
      ph.x = *reinterpret_cast<const uint32_t*>(&hi01); ph.y = *reinterpret_cast<const uint32_t*>(&hi23);
      pl.x = *reinterpret_cast<const uint32_t*>(&lo01); pl.y = *reinterpret_cast<const uint32_t*>(&lo23);
      *reinterpret_cast<uint2*>(o_hi + r[u] * ldo + n) = ph;
      *reinterpret_cast<uint2*>(o_lo + r[u] * ldo + n) = pl;
    }
  }
  if (vmax > 60000.f) atomicOr(overflow, 1);
}

// ---- weight preparation: per-matrix power-of-two scale, (hi, lo) split ----
constexpr int MAX_SEG = 16;
struct Seg {
  const float* src;
  int64_t ld;
  int rows, cols, mat, row0;
};
struct PrepArgs {
  Seg seg[MAX_SEG];
  int n_seg;
  unsigned int* amax_bits;      // [n_mat]
  float* unscale;               // [n_mat]
  __half* hi[MAX_SEG];          // per matrix
  __half* lo[MAX_SEG];
  __half* hiT[MAX_SEG];         // optional transposed copy (backward data-gradient GEMMs), same scale
  __half* loT[MAX_SEG];
  int64_t ldo[MAX_SEG];
  // side tables built by the same launch
  const float* b0;
  const float* b1;
  float* bias_cat;
  const float* w_init;
  float* wet;
  int H, fa, fb;
};

// blockIdx.y < n_seg: amax of one weight segment.  The two extra y values build the small fp32 side tables that do
// not depend on the scales: [b_i | b_o] and W_e^T (edge_init.weight[:, fa:] transposed).
__global__ void __launch_bounds__(256) prep_amax_kernel(const PrepArgs a) {
  if ((int)blockIdx.y == a.n_seg) {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < a.H; i += gridDim.x * 256) {
      a.bias_cat[i] = __ldg(a.b0 + i);
      a.bias_cat[a.H + i] = __ldg(a.b1 + i);
    }
    return;
  }
  if ((int)blockIdx.y == a.n_seg + 1) {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < a.fb * a.H; i += gridDim.x * 256) {
      const int k = i / a.H, n = i % a.H;
      a.wet[i] = __ldg(a.w_init + (int64_t)n * (a.fa + a.fb) + a.fa + k);
    }
    return;
  }
  const Seg s = a.seg[blockIdx.y];
  float m = 0.f;
  for (int r = blockIdx.x; r < s.rows; r += gridDim.x) {          // rows over blocks, columns over threads: coalesced
    const float* row = s.src + (int64_t)r * s.ld;
#pragma unroll 4
    for (int c = threadIdx.x; c < s.cols; c += 256) m = fmaxf(m, fabsf(__ldg(row + c)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
    atomicMax(a.amax_bits + s.mat, __float_as_uint(m));            // max is order-independent
  }
}

__device__ __forceinline__ float weight_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  return exp2f(13.f - ceilf(log2f(amax)));        // amax * scale in (2^12, 2^13]
}

// 32 x 32 element tiles: rows are read and written coalesced; the transposed copy goes through shared memory
__global__ void __launch_bounds__(256) prep_split_kernel(const PrepArgs a) {
  __shared__ __half th[32][34], tl[32][34];
  const Seg s = a.seg[blockIdx.y];
  const float sc = weight_scale(__uint_as_float(a.amax_bits[s.mat]));
  if (blockIdx.x == 0 && threadIdx.x == 0 && s.row0 == 0) a.unscale[s.mat] = 1.f / sc;
  __half* hi = a.hi[s.mat];
  __half* lo = a.lo[s.mat];
  __half* hiT = a.hiT[s.mat];
  __half* loT = a.loT[s.mat];
  const int64_t ldo = a.ldo[s.mat];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_c = (s.cols + 31) / 32, tiles = ((s.rows + 31) / 32) * tiles_c;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      __half h = __float2half_rn(0.f), l = h;
      if (r < s.rows && c < s.cols) {
        split_f16(__ldg(s.src + (int64_t)r * s.ld + c) * sc, h, l);
        hi[(s.row0 + r) * ldo + c] = h;
        lo[(s.row0 + r) * ldo + c] = l;
      }
      th[ty + 8 * i][tx] = h;
      tl[ty + 8 * i][tx] = l;
    }
    if (hiT) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;                // T[c][row0 + r] = M[r][c]
        if (r < s.rows && c < s.cols) {
          hiT[(int64_t)c * ldo + s.row0 + r] = th[tx][ty + 8 * i];
          loT[(int64_t)c * ldo + s.row0 + r] = tl[tx][ty + 8 * i];
        }
      }
      __syncthreads();
    }
  }
}

// greedy packing of consecutive whole reactions into 128-row tiles (bonds and atoms both <= 128).
// The greedy rule is a chain: the tile that starts at reaction s ends at next(s) = the largest j with
// bonds(s..j) <= 128 and atoms(s..j) <= 128, and the next tile starts there.  next() of every reaction is independent
// (a bisection over at most 64 followers: a reaction has >= 2 directed bonds), so one block computes it for a chunk of
// reactions in parallel from the offsets staged in shared memory (with a 64-reaction halo), ONE thread only follows the
// chain through shared memory (one dependent load per TILE instead of ~160 cycles per REACTION: 668 -> ~45 us at 8192
// reactions) and all threads write the chunk's tile records.  Same records as the host twin below (tc_plan_host).
constexpr int PLAN_CH = 2048, PLAN_HALO = TM / 2, PLAN_THREADS = 512;
__global__ void __launch_bounds__(PLAN_THREADS) tile_plan_kernel(const int32_t* __restrict__ in_ptr,
                                                                 const int32_t* __restrict__ atom_ptr, int64_t B,
                                                                 int32_t* __restrict__ tile_info,
                                                                 int32_t* __restrict__ status) {
  __shared__ int32_t a_s[PLAN_CH + PLAN_HALO + 1], e_s[PLAN_CH + PLAN_HALO + 1];   // (atom, bond) offsets of the chunk
  __shared__ uint16_t nxt_s[PLAN_CH];         // local index of the reaction behind the tile that starts here
  __shared__ uint16_t start_s[PLAN_CH];       // tile starts of the chunk, in chain order
  __shared__ int64_t s_carry;                 // global index of the next tile start
  __shared__ int n_start, t_base, ok_s;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) { s_carry = 0; t_base = 0; ok_s = 1; n_start = 0; }
  __syncthreads();
  for (int64_t g0 = 0; g0 < B; g0 += PLAN_CH) {
    const int n = (int)((B - g0) < PLAN_CH ? (B - g0) : PLAN_CH);
    const int m = (int)((B - g0) < n + PLAN_HALO ? (B - g0) : n + PLAN_HALO);
    for (int i = tid; i <= m; i += nt) {
      const int32_t a = __ldg(atom_ptr + g0 + i);
      a_s[i] = a;
      e_s[i] = __ldg(in_ptr + a);                  // bonds of a reaction are the in-bonds of its atoms
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      const int e0 = e_s[i], a0 = a_s[i];
      const int ne = e_s[i + 1] - e0, na = a_s[i + 1] - a0;
      if (ne > TM || na > TM || ne <= 0 || na <= 0 || (ne & 1)) ok_s = 0;      // not tileable: the plan is void
      int lo = i + 1, hi = i + PLAN_HALO < m ? i + PLAN_HALO : m;
      while (lo < hi) {                            // largest j in [i + 1, hi] whose span still fits a tile
        const int mid = (lo + hi + 1) >> 1;
        if (e_s[mid] - e0 <= TM && a_s[mid] - a0 <= TM) lo = mid; else hi = mid - 1;
      }
      nxt_s[i] = (uint16_t)lo;
    }
    __syncthreads();
    if (tid == 0) {
      int s = (int)(s_carry - g0), c = 0;
      while (s < n) { start_s[c++] = (uint16_t)s; s = nxt_s[s]; }
      n_start = c;
      s_carry = g0 + s;
    }
    __syncthreads();
    for (int c = tid; c < n_start; c += nt) {
      const int s = start_s[c], j = nxt_s[s];
      int32_t* ti = tile_info + (int64_t)(t_base + c) * 8;
      ti[0] = e_s[s]; ti[1] = e_s[j] - e_s[s]; ti[2] = a_s[s]; ti[3] = a_s[j] - a_s[s];
      ti[4] = (int32_t)(g0 + s); ti[5] = j - s;
    }
    __syncthreads();
    if (tid == 0) t_base += n_start;
    __syncthreads();
  }
  if (tid == 0) {
    status[0] = t_base;
    status[1] = ok_s;
  }
}

// every bond of a tile must have both endpoints among the tile's atoms (reactions are disjoint graphs)
__global__ void tile_check_kernel(const int32_t* __restrict__ tile_info, const int32_t* __restrict__ src,
                                  const int32_t* __restrict__ dst, int32_t* __restrict__ status) {
  const int tile = blockIdx.x;
  if (tile >= status[0]) return;       // launched over an upper bound of the tile count
  const int ebase = tile_info[tile * 8], ecount = tile_info[tile * 8 + 1];
  const int abase = tile_info[tile * 8 + 2], acount = tile_info[tile * 8 + 3];
  for (int j = threadIdx.x; j < ecount; j += blockDim.x) {
    const int s = src[ebase + j] - abase, d = dst[ebase + j] - abase;
    if (s < 0 || s >= acount || d < 0 || d >= acount) status[1] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// Tensor maps are pure functions of (pointer, shape, stride, box, type, swizzle).  The allocator hands the same
// buffers back step after step, so the encoded descriptors are cached per host thread (cuTensorMapEncodeTiled costs
// about a microsecond and a forward needs ~30 of them).
struct MapKey {
  const void* p;
  uint64_t d0, d1, stride;
  uint32_t b0, b1;
  int dtype, swizzle;
  bool operator==(const MapKey& o) const {
    return p == o.p && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1 && dtype == o.dtype &&
           swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.p * 0x9E3779B97F4A7C15ull;
    h ^= (k.d0 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.d1 * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2));
    h ^= (k.stride + ((uint64_t)k.b0 << 32) + k.b1 + ((uint64_t)k.dtype << 20) + ((uint64_t)k.swizzle << 28) + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};
int encode_map_cached(CUtensorMap* tm, CUtensorMapDataType dtype, const void* basep, uint64_t d0, uint64_t d1,
                      uint64_t stride_bytes, uint32_t b0, uint32_t b1, CUtensorMapSwizzle swz) {
  thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static const bool use_cache = getenv("CGR_NO_MAP_CACHE") == nullptr;
  const MapKey key{basep, d0, d1, stride_bytes, b0, b1, (int)dtype, (int)swz};
  if (use_cache) {
    auto it = cache.find(key);
    if (it != cache.end()) { memcpy(tm, &it->second, sizeof(CUtensorMap)); return CGR_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { cgr_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CGR_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dtype, 2, (void*)basep, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { cgr_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return CGR_ERR_ARG; }
  if (use_cache) {
    if (cache.size() >= 1024) cache.clear();
    cache.emplace(key, *tm);
  }
  return CGR_OK;
}

// fp16 matrix [rows, cols] with row stride `ld` elements; box = [box_rows, 64 cols]; 128-byte swizzle;
// out-of-bounds elements (K tail, row tail) are zero-filled by the TMA unit.
int make_map(CUtensorMap* tm, const __half* basep, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return encode_map_cached(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, basep, (uint64_t)cols, (uint64_t)rows,
                           (uint64_t)ld * sizeof(__half), (uint32_t)BK, (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

// fp32 matrix [rows, cols], row stride `ld` elements, dense (unswizzled) box [box_rows, box_cols]
int make_map_f32(CUtensorMap* tm, const float* basep, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                 int box_rows) {
  return encode_map_cached(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, basep, (uint64_t)cols, (uint64_t)rows,
                           (uint64_t)ld * sizeof(float), (uint32_t)box_cols, (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_NONE);
}

int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

struct WLayout {                 // prepared-weight buffer: [amax | unscale | bias_cat | matrices (hi, lo)...]
  int n_mat;
  int64_t kp_x, kp_h;
  size_t off_amax, off_unscale, off_bias, off_wet, off_hi[MAX_SEG], off_lo[MAX_SEG], off_hiT[MAX_SEG], off_loT[MAX_SEG], total;
  int64_t rows[MAX_SEG], ld[MAX_SEG];
};

WLayout wlayout(const cgr_params_t* p) {
  WLayout w;
  const int64_t H = p->hidden;
  w.n_mat = p->depth + 2;
  w.kp_x = round_up(p->fa, BK);
  w.kp_h = round_up(H, BK);
  size_t off = 0;
  w.off_amax = off; off += cgr_align_up(MAX_SEG * sizeof(unsigned int), 256);
  w.off_unscale = off; off += cgr_align_up(MAX_SEG * sizeof(float), 256);
  w.off_bias = off; off += cgr_align_up(2 * H * sizeof(float), 256);
  w.off_wet = off; off += cgr_align_up((size_t)(p->fb > 0 ? p->fb : 1) * H * sizeof(float), 256);
  for (int m = 0; m < w.n_mat; ++m) {
    w.rows[m] = m == 0 ? 2 * H : H;
    w.ld[m] = m == 0 ? w.kp_x : w.kp_h;
    const size_t bytes = cgr_align_up((size_t)w.rows[m] * w.ld[m] * sizeof(__half), 1024);
    w.off_hi[m] = off; off += bytes;
    w.off_lo[m] = off; off += bytes;
    w.off_hiT[m] = w.off_loT[m] = 0;
    if (m >= 1) {                 // square [H, H] matrices: transposed copies for dh = dy W
      w.off_hiT[m] = off; off += bytes;
      w.off_loT[m] = off; off += bytes;
    }
  }
  w.total = off;
  return w;
}

template <int BN, int EPI, bool RELU, int NT>
int launch_gemm_t(const TcGemmParams& prm, int m_tiles, int n_slices, const char* name, bool pdl, cudaStream_t st) {
  using C = Cfg<BN, EPI, NT>;
  constexpr int THREADS = NT;
  static bool attr_done = false;      // benign race: the attribute is idempotent
  if (!attr_done) {
    CGR_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, EPI, RELU, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
    attr_done = true;
  }
  CgrRange prof(name, st);
  cgr_note_launch(name, st, 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  const int mc = (EPI == EPI_PLAIN && prm.mc > 1) ? prm.mc : 1;     // weight multicast: clusters of mc row tiles
  cfg.gridDim = dim3((unsigned)n_slices, (unsigned)(cgr_ceil_div(m_tiles, mc) * mc));   // (tiles past the end store nothing)
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {                                     // pdl: this launch may overlap the tail of its predecessor
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (mc > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = (unsigned)mc;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CGR_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, EPI, RELU, NT>, prm));
  return CGR_OK;
}

constexpr int BN_SMALL = 80, BN_LARGE = 208;
constexpr int NT_SMALL = 256;     // narrow slices run two CTAs per SM: one CTA's epilogue overlaps the other's MMAs

// Slice width: wide slices (208) amortise the A-operand fetch of the SS-mode MMA and halve the A re-reads;
// narrow slices (80) give small batches enough CTAs to occupy the 148 SMs.
int choose_bn(int64_t m_tiles, int64_t n_total) {
  static const int forced = getenv("CGR_BN") ? atoi(getenv("CGR_BN")) : 0;     // experiments: force a slice width
  if (forced == BN_LARGE || forced == BN_SMALL) return forced;
  return m_tiles * cgr_ceil_div(n_total, BN_LARGE) >= 148 ? BN_LARGE : BN_SMALL;
}
int chunk_cols(int bn) { return bn > 128 ? bn / 2 : bn; }

// two_per_sm: 256-thread CTAs, two per SM (throughput); otherwise 512-thread CTAs, one per SM (latency)
template <int EPI>
int launch_gemm(const TcGemmParams& prm, int bn, int m_tiles, bool relu, const char* name, bool pdl, bool two_per_sm,
                cudaStream_t st) {
  const int n_slices = (int)cgr_ceil_div(prm.n_total, bn);
  if (bn == BN_SMALL && !two_per_sm) {
    return relu ? launch_gemm_t<BN_SMALL, EPI, true, 512>(prm, m_tiles, n_slices, name, pdl, st)
                : launch_gemm_t<BN_SMALL, EPI, false, 512>(prm, m_tiles, n_slices, name, pdl, st);
  }
  if (bn == BN_LARGE) {
    return relu ? launch_gemm_t<BN_LARGE, EPI, true, 512>(prm, m_tiles, n_slices, name, pdl, st)
                : launch_gemm_t<BN_LARGE, EPI, false, 512>(prm, m_tiles, n_slices, name, pdl, st);
  }
  return relu ? launch_gemm_t<BN_SMALL, EPI, true, NT_SMALL>(prm, m_tiles, n_slices, name, pdl, st)
              : launch_gemm_t<BN_SMALL, EPI, false, NT_SMALL>(prm, m_tiles, n_slices, name, pdl, st);
}

// ---- persistent atom projection (tc_proj.cuh): one CTA per SM walks the (row tile, column slice) units ----
constexpr int PROJ_BN_3STAGE = 160, PROJ_BN_WIDE = 208;
// slice width of the persistent kernel: 160 (three operand stages) unless 208 pads the output width less; 0 = use
// the one-unit-per-CTA kernel
int proj_bn(int bn_default, int n_total) {
  static const bool off = getenv("CGR_AP_OLD") != nullptr;       // experiments: the one-unit-per-CTA kernel
  static const int forced = getenv("CGR_AP_BN") ? atoi(getenv("CGR_AP_BN")) : 0;
  if (off || bn_default != BN_LARGE || n_total % 8 != 0) return 0;
  if (forced == PROJ_BN_3STAGE || forced == PROJ_BN_WIDE) return forced;
  const int64_t pad160 = cgr_ceil_div(n_total, PROJ_BN_3STAGE) * PROJ_BN_3STAGE - n_total;
  const int64_t pad208 = cgr_ceil_div(n_total, PROJ_BN_WIDE) * PROJ_BN_WIDE - n_total;
  return pad160 <= pad208 ? PROJ_BN_3STAGE : PROJ_BN_WIDE;
}
// weight multicast of the persistent projection: clusters of `mc` CTAs (consecutive row tiles, same slice) share every
// weight chunk.  0 / 1 = off.
int proj_mc(int bn) {
  // measured (group of 20 cfg-2 batches): 104.3 us without, 104.7 us with clusters of 2, 116.6 us with clusters of 4 --
  // L2 already merges the same line requested by a few SMs at about the same time, so multicast at this cluster size
  // saves no L2 throughput (the kernel runs at ~75-80 % of the ~6.3 KB/clk the L2 delivers to the SMs); off
  static const int mc_env = getenv("CGR_AP_MC") ? atoi(getenv("CGR_AP_MC")) : 1;
  const int mc = (mc_env == 2 || mc_env == 4) ? mc_env : 1;
  return (mc > 1 && bn % (8 * mc) == 0) ? mc : 1;            // shares of whole 8-row swizzle groups only
}
bool proj_cg2_wanted(int bn, int n_total);
int proj_multicast(TcGemmParams* ap, const __half* w_hi, const __half* w_lo, int64_t rows, int64_t cols, int64_t ld, int bn) {
  if (proj_cg2_wanted(bn, (int)rows)) {          // CTA-pair kernel: each CTA stages half of the slice's weight rows
    int rc;
    if ((rc = make_map(&ap->tmB_hi_mc, w_hi, rows, cols, ld, bn / 2))) return rc;
    if ((rc = make_map(&ap->tmB_lo_mc, w_lo, rows, cols, ld, bn / 2))) return rc;
    ap->mc = -2;
    return CGR_OK;
  }
  const int mc = proj_mc(bn);
  if (mc == 1) return CGR_OK;
  int rc;
  if ((rc = make_map(&ap->tmB_hi_mc, w_hi, rows, cols, ld, bn / mc))) return rc;
  if ((rc = make_map(&ap->tmB_lo_mc, w_lo, rows, cols, ld, bn / mc))) return rc;
  ap->mc = mc;
  return CGR_OK;
}
template <int BN>
int launch_proj_t(const TcGemmParams& prm, int m_tiles, cudaStream_t st) {
  using C = tcp::PCfg<BN>;
  const int mc = prm.mc > 1 ? prm.mc : 1;
  static bool attr_done = false;      // benign race: the attribute is idempotent
  static int n_sm = 0, max_clusters[5] = {0, 0, 0, 0, 0};
  if (!attr_done) {
    CGR_CUDA(cudaFuncSetAttribute(tcp::tc_proj_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    int dev = 0;
    CGR_CUDA(cudaGetDevice(&dev));
    CGR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(tcp::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)mc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mc > 1 ? 1 : 0;
  if (mc > 1 && max_clusters[mc] == 0) {
    // a persistent grid must be co-resident: clusters the device can hold at once (GPCs with an odd number of SMs
    // leave one out)
    cfg.gridDim = dim3((unsigned)(n_sm / mc * mc));
    int n = 0;
    CGR_CUDA(cudaOccupancyMaxActiveClusters(&n, tcp::tc_proj_kernel<BN>, &cfg));
    max_clusters[mc] = n > 0 ? n : 1;
  }
  const int n_slices = (int)cgr_ceil_div(prm.n_total, BN);
  const int n_units = (int)cgr_ceil_div(m_tiles, mc) * n_slices;         // cluster units
  const int slots = mc > 1 ? max_clusters[mc] : n_sm;
  CgrRange prof("tc_atom_proj", st);
  cgr_note_launch("tc_atom_proj", st, 1);
  cfg.gridDim = dim3((unsigned)((n_units < slots ? n_units : slots) * mc));
  CGR_CUDA(cudaLaunchKernelEx(&cfg, tcp::tc_proj_kernel<BN>, prm, n_units, n_slices));
  return CGR_OK;
}
// CTA-pair variant (tcgen05 cta_group::2): 160-wide slices only; needs the half-height weight maps (tmB_*_mc, mc = 2).
// Measured (group of 20 cfg-2 batches): 109.5 us against 100.4 us for the one-CTA kernel on the same box -- halving the
// weight rows a CTA reads from shared memory per MMA does not speed the MMAs up, and the peer's "chunk landed" relay
// lengthens every stage's round trip; optional (CGR_AP_CG2=1), parity-tested, off.  Slice widths measured under ncu
// (cold, ~1 GHz): 128 / 160 / 208 / 256 columns = 235 / 192 / 178 / 203 k cycles for the same launch (256 pads 2H = 800
// to 1024).
bool proj_cg2_wanted(int bn, int n_total) {
  static const bool on = getenv("CGR_AP_CG2") != nullptr;
  return on && bn == PROJ_BN_3STAGE && n_total % PROJ_BN_3STAGE == 0;
}
int launch_proj_cg2(const TcGemmParams& prm, int m_tiles, cudaStream_t st) {
  using C = tcp::P2Cfg<PROJ_BN_3STAGE>;
  static bool attr_done = false;
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(tcp::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!attr_done) {
    CGR_CUDA(cudaFuncSetAttribute(tcp::tc_proj_cg2_kernel<PROJ_BN_3STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    int dev = 0, n_sm = 0;
    CGR_CUDA(cudaGetDevice(&dev));
    CGR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    cfg.gridDim = dim3((unsigned)(n_sm / 2 * 2));
    int n = 0;
    CGR_CUDA(cudaOccupancyMaxActiveClusters(&n, tcp::tc_proj_cg2_kernel<PROJ_BN_3STAGE>, &cfg));
    max_clusters = n > 0 ? n : 1;
    attr_done = true;
  }
  const int n_slices = (int)cgr_ceil_div(prm.n_total, PROJ_BN_3STAGE);
  const int n_units = (int)cgr_ceil_div(m_tiles, 2) * n_slices;
  CgrRange prof("tc_atom_proj", st);
  cgr_note_launch("tc_atom_proj", st, 1);
  cfg.gridDim = dim3((unsigned)((n_units < max_clusters ? n_units : max_clusters) * 2));
  CGR_CUDA(cudaLaunchKernelEx(&cfg, tcp::tc_proj_cg2_kernel<PROJ_BN_3STAGE>, prm, n_units, n_slices));
  return CGR_OK;
}
int launch_proj(const TcGemmParams& prm, int bn, int m_tiles, cudaStream_t st) {
  if (prm.mc == -2) return launch_proj_cg2(prm, m_tiles, st);
  return bn == PROJ_BN_3STAGE ? launch_proj_t<PROJ_BN_3STAGE>(prm, m_tiles, st) : launch_proj_t<PROJ_BN_WIDE>(prm, m_tiles, st);
}

// ---- fused forward: all bond layers + readout of a tile group in one cluster launch (tc_fwd.cuh) ----
constexpr int FWD_BN_WIDE = 208, FWD_BN_NARROW = 80;

template <int BN, bool RELU>
int launch_fwd_t(const tcf::FwdParams& prm, int n_groups, int S, bool pdl, cudaStream_t st) {
  using C = tcf::FCfg<BN>;
  static bool attr_done = false;      // benign race: the attribute is idempotent
  if (!attr_done) {
    CGR_CUDA(cudaFuncSetAttribute(tcf::tc_fwd_kernel<BN, RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  CgrRange prof("tc_fwd_fused", st);
  cgr_note_launch("tc_fwd_fused", st, 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(n_groups * S));
  cfg.blockDim = dim3(tcf::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  CGR_CUDA(cudaLaunchKernelEx(&cfg, tcf::tc_fwd_kernel<BN, RELU>, prm));
  return CGR_OK;
}

struct FwdChoice { int bn, S, tpc; };
// wide slices + two tiles per cluster (the epilogue of one tile hides behind the other's MMAs): throughput;
// narrow slices + one tile per cluster (most CTAs per tile): latency of a lone small batch
bool choose_fwd(int64_t T, int H, bool throughput, FwdChoice* c) {
  const int s_wide = (int)cgr_ceil_div(H, FWD_BN_WIDE), s_narrow = (int)cgr_ceil_div(H, FWD_BN_NARROW);
  static const char* forced = getenv("CGR_FWD_CFG");      // experiments: wide1 | wide2 | narrow1 | narrow2
  if (forced) {
    const bool wide = strncmp(forced, "wide", 4) == 0;
    c->bn = wide ? FWD_BN_WIDE : FWD_BN_NARROW;
    c->S = wide ? s_wide : s_narrow;
    c->tpc = forced[strlen(forced) - 1] == '2' ? 2 : 1;
    return c->S <= 8;
  }
  if (!throughput && s_narrow <= 8 && T * s_narrow <= 120) { *c = FwdChoice{FWD_BN_NARROW, s_narrow, 1}; return true; }
  if (s_wide <= 8) { *c = FwdChoice{FWD_BN_WIDE, s_wide, T >= 2 ? 2 : 1}; return true; }
  return false;
}

}  // namespace

// ------------------------------------------------------------------------------------------------

int tc_split_features(const float* x, int64_t n, int fa, void* x_hi, void* x_lo, int* status, cudaStream_t st) {
  CGR_CHECK_ARG(x && x_hi && x_lo && status && n >= 0 && fa > 0, "tc_split_features: bad argument");
  if (n == 0) return CGR_OK;
  CgrRange prof("tc_split_x", st);
  cgr_note_launch("tc_split_x", st, 1);
  split_rows_kernel<<<(unsigned)cgr_ceil_div(n, 8), 256, 0, st>>>(x, fa, n, fa, (__half*)x_hi, (__half*)x_lo,
                                                                  round_up(fa, BK), status, FLAG_X);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int tc_flag_begin(int* flag, cudaStream_t st) {
  if (!flag) return CGR_OK;
  cgr_note_launch("tc_flag", st, 1);
  flag_begin_kernel<<<1, 1, 0, st>>>(flag);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}
int tc_poison_outputs(float* out, int64_t n, const int* flag, cudaStream_t st) {
  if (!flag || n <= 0) return CGR_OK;
  cgr_note_launch("tc_flag", st, 1);
  poison_kernel<<<(unsigned)(n < 65536 ? cgr_ceil_div(n, 256) : 256), 256, 0, st>>>(out, n, flag);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

static long long* g_tc_dbg = nullptr;   // debug: phase time stamps of the bond-layer kernel
void tc_set_debug_buffer(long long* p) { g_tc_dbg = p; }

size_t tc_weights_bytes(const cgr_params_t* p) { return wlayout(p).total; }

int tc_prepare_weights(const cgr_params_t* p, void* wbuf, size_t wbuf_bytes, cudaStream_t st) {
  CGR_CHECK_ARG(p->depth + 3 <= MAX_SEG, "tcgen05 engine supports depth <= %d", MAX_SEG - 3);
  const WLayout w = wlayout(p);
  CGR_CHECK_ARG(wbuf && wbuf_bytes >= w.total, "tc_prepare_weights: buffer too small");
  char* b = (char*)wbuf;
  const int H = p->hidden, fa = p->fa, fb = p->fb;
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  a.amax_bits = (unsigned int*)(b + w.off_amax);
  a.unscale = (float*)(b + w.off_unscale);
  for (int m = 0; m < w.n_mat; ++m) {
    a.hi[m] = (__half*)(b + w.off_hi[m]);
    a.lo[m] = (__half*)(b + w.off_lo[m]);
    a.hiT[m] = m >= 1 ? (__half*)(b + w.off_hiT[m]) : nullptr;
    a.loT[m] = m >= 1 ? (__half*)(b + w.off_loT[m]) : nullptr;
    a.ldo[m] = w.ld[m];
  }
  int ns = 0;
  a.seg[ns++] = Seg{p->w_init, fa + fb, H, fa, 0, 0};              // W_x   = edge_init.weight[:, :Fa]
  a.seg[ns++] = Seg{p->w_e2n, fa + H, H, fa, 0, H};                // W_ox  = edge_to_node.weight[:, :Fa]
  for (int l = 0; l < p->depth; ++l) a.seg[ns++] = Seg{p->w_conv[l], H, H, H, 1 + l, 0};
  a.seg[ns++] = Seg{p->w_e2n + fa, fa + H, H, H, p->depth + 1, 0};  // W_os  = edge_to_node.weight[:, Fa:]
  a.n_seg = ns;
  CgrRange prof("tc_prep_weights", st);
  cgr_note_launch("tc_prep_weights", st, 2);
  a.b0 = p->b_init; a.b1 = p->b_e2n; a.bias_cat = (float*)(b + w.off_bias);
  a.w_init = p->w_init; a.wet = (float*)(b + w.off_wet);
  a.H = H; a.fa = fa; a.fb = fb;
  CGR_CUDA(cudaMemsetAsync(a.amax_bits, 0, MAX_SEG * sizeof(unsigned int), st));
  prep_amax_kernel<<<dim3(200, (unsigned)ns + 2), 256, 0, st>>>(a);
  prep_split_kernel<<<dim3(256, (unsigned)ns), 256, 0, st>>>(a);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int tc_plan_build(const int32_t* in_ptr, const int32_t* atom_ptr, const int32_t* src, const int32_t* dst,
                  int64_t n_rxn, int32_t* tile_info, int32_t* status, cudaStream_t st) {
  CGR_CHECK_ARG(in_ptr && atom_ptr && tile_info && status && n_rxn > 0, "tc_plan_build: bad argument");
  (void)src; (void)dst;            // endpoints are validated by tc_plan_check once the tile count is known
  cgr_note_launch("tc_plan", st, 1);
  tile_plan_kernel<<<1, PLAN_THREADS, 0, st>>>(in_ptr, atom_ptr, n_rxn, tile_info, status);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

// Host-side twin of tile_plan_kernel for callers that already know the per-reaction offsets (reaction store, host
// entry): same greedy rule, no device work, no synchronisation.  tile_info is a HOST array of [n_rxn][8] ints.
int tc_plan_host(const int64_t* atom_ptr, const int64_t* edge_ptr, int64_t n_rxn, int32_t* tile_info, int64_t* n_tiles) {
  CGR_CHECK_ARG(atom_ptr && edge_ptr && tile_info && n_tiles && n_rxn > 0, "tc_plan_host: bad argument");
  int64_t t = -1;
  int used_e = TM + 1, used_a = TM + 1;
  for (int64_t g = 0; g < n_rxn; ++g) {
    const int64_t ne = edge_ptr[g + 1] - edge_ptr[g], na = atom_ptr[g + 1] - atom_ptr[g];
    if (ne > TM || na > TM || ne <= 0 || na <= 0 || (ne & 1)) {
      cgr_set_error("reaction %lld has %lld bonds / %lld atoms: not tileable for the tcgen05 engine", (long long)g,
                    (long long)ne, (long long)na);
      return CGR_ERR_UNSUPPORTED;
    }
    if (used_e + ne > TM || used_a + na > TM) {
      ++t;
      int32_t* ti = tile_info + t * 8;
      ti[0] = (int32_t)edge_ptr[g]; ti[1] = 0; ti[2] = (int32_t)atom_ptr[g]; ti[3] = 0; ti[4] = (int32_t)g; ti[5] = 0;
      ti[6] = 0; ti[7] = 0;
      used_e = 0; used_a = 0;
    }
    used_e += (int)ne; used_a += (int)na;
    int32_t* ti = tile_info + t * 8;
    ti[1] = used_e; ti[3] = used_a; ti[5] += 1;
  }
  *n_tiles = t + 1;
  return CGR_OK;
}

int tc_plan_check(const int32_t* tile_info, int64_t n_tiles, const int32_t* src, const int32_t* dst, int32_t* status,
                  cudaStream_t st) {
  if (n_tiles <= 0) return CGR_OK;
  cgr_note_launch("tc_plan", st, 1);
  tile_check_kernel<<<(unsigned)n_tiles, 128, 0, st>>>(tile_info, src, dst, status);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

namespace {
struct TcWs {
  size_t off_w, off_xhi, off_xlo, off_pq, off_h0, off_hhi[2], off_hlo[2], off_partial, off_flag, total;
  int64_t kp_x, kp_h, rows_pad;
  int n_slices;
};
TcWs tc_ws(const cgr_params_t* p, const cgr_graph_t* g, bool need_w) {
  const bool need_x = !(g->x_hi && g->x_lo);
  TcWs w;
  const int64_t H = p->hidden, N = g->n_atoms, T = g->n_tiles, B = g->n_rxn;
  w.kp_x = round_up(p->fa, BK);
  w.kp_h = round_up(H, BK);
  w.rows_pad = T * TM;
  w.n_slices = (int)cgr_ceil_div(H, BN_SMALL);   // upper bound over both slice widths
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += cgr_align_up(bytes, 1024); return o; };
  w.off_w = take(need_w ? tc_weights_bytes(p) : 0);
  w.off_xhi = take(need_x ? (size_t)N * w.kp_x * sizeof(__half) : 0);
  w.off_xlo = take(need_x ? (size_t)N * w.kp_x * sizeof(__half) : 0);
  w.off_pq = take((size_t)N * 2 * H * sizeof(float));
  w.off_h0 = take((size_t)w.rows_pad * H * sizeof(float));
  for (int i = 0; i < 2; ++i) {
    w.off_hhi[i] = take((size_t)w.rows_pad * w.kp_h * sizeof(__half));
    w.off_hlo[i] = take((size_t)w.rows_pad * w.kp_h * sizeof(__half));
  }
  w.off_partial = take((size_t)w.n_slices * B * sizeof(float));
  w.off_flag = take(256);
  w.total = off + 1024;
  return w;
}
}  // namespace

namespace {
// Activations the fused training forward keeps for the fused backward (the cgr_saved_t.tc_blob buffer):
// every layer's output as tile-packed FP16 (hi, lo) rows -- the backward's GEMM operands and ReLU masks --
// then h_0 in fp32 (skip-weight gradient) and hv in fp32 (mask of the readout backward).
struct TcSavedLayout {
  int64_t kp_h, rows_pad;
  size_t off_hhi[MAX_SEG], off_hlo[MAX_SEG], hl_bytes, off_h0, off_hv, off_w, total;
};
TcSavedLayout tc_saved_layout(const cgr_params_t* p, const cgr_graph_t* g) {
  TcSavedLayout L;
  const int64_t H = p->hidden;
  L.kp_h = round_up(H, BK);
  L.rows_pad = g->n_tiles * TM;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += cgr_align_up(bytes, 1024); return o; };
  for (int l = 0; l <= p->depth; ++l) {
    L.off_hhi[l] = take((size_t)L.rows_pad * L.kp_h * sizeof(__half));
    L.off_hlo[l] = take((size_t)L.rows_pad * L.kp_h * sizeof(__half));
  }
  L.hl_bytes = off;
  L.off_h0 = take((size_t)L.rows_pad * H * sizeof(float));
  L.off_hv = take((size_t)g->n_atoms * H * sizeof(float));
  L.off_w = take(tc_weights_bytes(p));              // prepared weights of this step (when the caller passed none)
  L.total = off + 1024;
  return L;
}
}  // namespace

bool tc_fused_training_ok(const cgr_params_t* p, const cgr_graph_t* g) {
  return g->tile_info && g->n_tiles > 0 && p->act == CGR_ACT_RELU && p->hidden % 4 == 0 && p->hidden <= 1024 &&
         p->fb <= 32 && p->depth + 4 <= MAX_SEG;
}
size_t tc_saved_bytes(const cgr_params_t* p, const cgr_graph_t* g) {
  if (!tc_fused_training_ok(p, g)) return 0;
  return tc_saved_layout(p, g).total;
}

size_t tc_forward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int training) {
  (void)training;
  if (!g->tile_info || g->n_tiles <= 0) return 0;
  return tc_ws(p, g, p->tc_weights == nullptr).total;
}

namespace {
// atom projection with wide slices: clusters of `mc` row tiles share every weight chunk (each CTA loads 1/mc of its rows
// and multicasts them) -- the kernel is operand-fill bound, the weights are 62 % of its fill
int ap_multicast(TcGemmParams* ap, const __half* w_hi, const __half* w_lo, int64_t rows, int64_t cols, int64_t ld, int bn) {
  static const int mc_env = getenv("CGR_AP_MC") ? atoi(getenv("CGR_AP_MC")) : 1;   // measured: no gain (share 0.22 vs 0.21), off
  const int mc = (mc_env == 2 || mc_env == 4) ? mc_env : 1;
  if (mc == 1 || bn % (8 * mc) != 0) return CGR_OK;          // shares of whole 8-row swizzle groups only
  int rc;
  if ((rc = make_map(&ap->tmB_hi_mc, w_hi, rows, cols, ld, bn / mc))) return rc;
  if ((rc = make_map(&ap->tmB_lo_mc, w_lo, rows, cols, ld, bn / mc))) return rc;
  ap->mc = mc;
  return CGR_OK;
}
}  // namespace

namespace {
// fused forward kernel: what every batch of a launch shares (weights, biases, shapes)
int fwd_fill_shared(tcf::FwdParams* prm, const cgr_params_t* p, const char* wbuf, const WLayout& wl, const FwdChoice& fc,
                    int64_t kp_h, bool fused_init, int fast) {
  const int H = p->hidden, d = p->depth;
  int rc;
  for (int l = 0; l <= d; ++l) {
    if ((rc = make_map(&prm->tmB_hi[l], (const __half*)(wbuf + wl.off_hi[1 + l]), H, H, wl.ld[1 + l], fc.bn))) return rc;
    if ((rc = make_map(&prm->tmB_lo[l], (const __half*)(wbuf + wl.off_lo[1 + l]), H, H, wl.ld[1 + l], fc.bn))) return rc;
    if (l < d) {
      prm->bias[l] = p->b_conv[l];
      prm->skip[l] = p->use_skip ? p->skip[l] : nullptr;
    }
  }
  prm->ldo = kp_h;
  prm->unscale = (const float*)(wbuf + wl.off_unscale);
  prm->wet = (const float*)(wbuf + wl.off_wet);
  prm->fb = p->fb;
  prm->fuse_init = fused_init ? 1 : 0;
  prm->w_ffn = p->w_ffn; prm->b_ffn = p->b_ffn;
  prm->depth = d; prm->H = H; prm->num_k = (int)cgr_ceil_div(H, BK); prm->act = p->act;
  prm->tiles_per_cluster = fc.tpc;
  prm->fast = fast;
  prm->dbg = g_tc_dbg;
  static const int publish = getenv("CGR_PUBLISH") ? atoi(getenv("CGR_PUBLISH")) : 2;     // experiments: 0 .. 3
  prm->publish_mode = publish;
  return CGR_OK;
}
// ... and one batch's part: operand buffers of its workspace, index arrays, outputs
int fwd_fill_batch(tcf::FwdBatch* bt, const cgr_graph_t* g, char* ws, const TcWs& w, int H, float* out) {
  int rc;
  __half* h_hi[2] = {(__half*)(ws + w.off_hhi[0]), (__half*)(ws + w.off_hhi[1])};
  __half* h_lo[2] = {(__half*)(ws + w.off_hlo[0]), (__half*)(ws + w.off_hlo[1])};
  for (int b = 0; b < 2; ++b) {
    if ((rc = make_map(&bt->tmA_hi[b], h_hi[b], w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&bt->tmA_lo[b], h_lo[b], w.rows_pad, H, w.kp_h, TM))) return rc;
    bt->o_hi[b] = h_hi[b];
    CGR_CHECK_ARG((char*)h_lo[b] - (char*)h_hi[b] == (char*)h_lo[0] - (char*)h_hi[0],
                  "tc forward: operand buffers are not laid out pairwise");
  }
  bt->lo_delta = (int64_t)((char*)h_lo[0] - (char*)h_hi[0]);
  bt->h0 = (float*)(ws + w.off_h0);
  bt->ea = g->edge_attr;
  bt->PQ = (const float*)(ws + w.off_pq);
  bt->tile_info = g->tile_info;
  bt->in_ptr = g->in_ptr; bt->in_idx = g->in_idx; bt->src = g->src; bt->atom_ptr = g->atom_ptr;
  bt->partial_out = (float*)(ws + w.off_partial);
  bt->out = out;
  bt->tile_counter = g->tc_status + 1;
  bt->overflow = g->tc_status;
  bt->n_rxn = g->n_rxn;
  bt->n_tiles = (int)g->n_tiles;
  bt->group0 = 0;
  bt->train_rows = 0;
  bt->hv_out = nullptr;
  return CGR_OK;
}
// training forward: the layers' operand pairs live in the saved blob, back to back (TcSavedLayout), and stay there
int fwd_fill_training(tcf::FwdBatch* bt, const cgr_params_t* p, char* blob, const TcSavedLayout& SL, int H) {
  const int64_t lo_delta = (int64_t)(SL.off_hlo[0] - SL.off_hhi[0]);
  for (int l = 0; l <= p->depth; ++l)
    CGR_CHECK_ARG((int64_t)SL.off_hhi[l] == (int64_t)SL.off_hhi[0] + 2 * l * lo_delta &&
                      (int64_t)SL.off_hlo[l] == (int64_t)SL.off_hhi[l] + lo_delta &&
                      lo_delta == SL.rows_pad * SL.kp_h * (int64_t)sizeof(__half),
                  "tc training forward: saved operand pairs are not contiguous");
  int rc;
  __half* base = (__half*)(blob + SL.off_hhi[0]);
  const int64_t rows = 2 * (int64_t)(p->depth + 1) * SL.rows_pad;
  CGR_CHECK_ARG(rows < (int64_t)1 << 31, "tc training forward: batch too large for one operand map");
  if ((rc = make_map(&bt->tmA_hi[0], base, rows, H, SL.kp_h, TM))) return rc;
  bt->o_hi[0] = base;
  bt->lo_delta = lo_delta;
  bt->h0 = (float*)(blob + SL.off_h0);
  bt->hv_out = (float*)(blob + SL.off_hv);
  bt->train_rows = SL.rows_pad;
  return CGR_OK;
}
int launch_fwd(const tcf::FwdParams& prm, const FwdChoice& fc, bool relu, int n_groups, bool pdl, cudaStream_t st) {
  if (fc.bn == FWD_BN_WIDE)
    return relu ? launch_fwd_t<FWD_BN_WIDE, true>(prm, n_groups, fc.S, pdl, st)
                : launch_fwd_t<FWD_BN_WIDE, false>(prm, n_groups, fc.S, pdl, st);
  return relu ? launch_fwd_t<FWD_BN_NARROW, true>(prm, n_groups, fc.S, pdl, st)
              : launch_fwd_t<FWD_BN_NARROW, false>(prm, n_groups, fc.S, pdl, st);
}
}  // namespace

int tc_gnn_forward(const cgr_params_t* p, const cgr_graph_t* g, float* out, cgr_saved_t* saved, int training,
                   uint64_t seed, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CGR_CHECK_ARG(g->tile_info && g->n_tiles > 0, "tcgen05 engine needs a tile plan (reactions of <= 128 bonds)");
  CGR_CHECK_ARG(g->n_tiles <= 65535, "tcgen05 engine: %lld row tiles exceed one launch (65535): split the batch "
                "(about 200k T1x-sized reactions per call)", (long long)g->n_tiles);
  char* blob = nullptr;               // training: per-layer activations are kept in the caller's blob
  TcSavedLayout SL;
  memset(&SL, 0, sizeof(SL));
  if (saved) {
    CGR_CHECK_ARG(saved->tc_blob && tc_fused_training_ok(p, g), "tcgen05 fused training forward: unsupported configuration or missing tc_blob");
    SL = tc_saved_layout(p, g);
    CGR_CHECK_ARG(saved->tc_blob_bytes >= SL.total, "tcgen05 fused training forward: tc_blob too small");
    blob = (char*)(((uintptr_t)saved->tc_blob + 1023) & ~(uintptr_t)1023);
    // pad rows of the tile-packed operands are read by the weight-gradient GEMMs (K = rows): they must be zero
    CGR_CUDA(cudaMemsetAsync(blob, 0, SL.hl_bytes, st));
  }
  CGR_CHECK_ARG(p->hidden % 4 == 0, "tcgen05 engine needs a hidden size that is a multiple of 4");
  CGR_CHECK_ARG(p->depth + 3 <= MAX_SEG, "tcgen05 engine supports depth <= %d", MAX_SEG - 3);
  CGR_CHECK_ARG(p->fb <= 32, "tcgen05 engine supports at most 32 bond features");
  CGR_CHECK_ARG(g->tc_status, "tcgen05 engine needs cgr_graph_t.tc_status ([1 + n_tiles] zero-initialised ints)");
  const bool need_w = p->tc_weights == nullptr;
  const TcWs w = tc_ws(p, g, need_w);
  CGR_CHECK_ARG(workspace && workspace_bytes >= w.total, "tc_gnn_forward: workspace too small");
  char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const int H = p->hidden, fa = p->fa, fb = p->fb, d = p->depth;
  const int64_t N = g->n_atoms, T = g->n_tiles, B = g->n_rxn;
  int rc;

  char* wbuf = need_w ? (blob ? blob + SL.off_w : ws + w.off_w) : (char*)p->tc_weights;
  if (need_w) {
    rc = tc_prepare_weights(p, wbuf, tc_weights_bytes(p), st);
    if (rc) return rc;
  }
  const WLayout wl = wlayout(p);
  const float* unscale = (const float*)(wbuf + wl.off_unscale);
  const float* bias_cat = (const float*)(wbuf + wl.off_bias);
  auto w_hi = [&](int m) { return (const __half*)(wbuf + wl.off_hi[m]); };
  auto w_lo = [&](int m) { return (const __half*)(wbuf + wl.off_lo[m]); };

  const bool need_x = !(g->x_hi && g->x_lo);
  __half* x_hi = need_x ? (__half*)(ws + w.off_xhi) : (__half*)g->x_hi;
  __half* x_lo = need_x ? (__half*)(ws + w.off_xlo) : (__half*)g->x_lo;
  float* PQ = (float*)(ws + w.off_pq);
  float* h0 = blob ? (float*)(blob + SL.off_h0) : (float*)(ws + w.off_h0);
  __half* h_hi[2] = {(__half*)(ws + w.off_hhi[0]), (__half*)(ws + w.off_hhi[1])};
  __half* h_lo[2] = {(__half*)(ws + w.off_hlo[0]), (__half*)(ws + w.off_hlo[1])};
  // operand buffers of layer l's input / output: ping-pong for inference, one pair per layer when saving
  auto hbuf_hi = [&](int l) { return blob ? (__half*)(blob + SL.off_hhi[l]) : h_hi[l & 1]; };
  auto hbuf_lo = [&](int l) { return blob ? (__half*)(blob + SL.off_hlo[l]) : h_lo[l & 1]; };
  float* partial = (float*)(ws + w.off_partial);
  static const bool use_pdl = getenv("CGR_NO_PDL") == nullptr;   // programmatic dependent launch between the kernels
  // grids larger than the machine, or forwards pipelined over streams by the caller, run two CTAs per SM
  auto two_per_sm = [&](int64_t ctas) { return p->tc_throughput != 0 || ctas > 148; };
  int* flag = g->tc_status;              // [0] fp16-range flag bits, [1..T] readout arrival counters
  int* tile_counter = g->tc_status + 1;
  const int fast = (p->tc_fast && !blob) ? 1 : 0;      // single-pass fp16: inference only (training keeps the parity mode)

  // 1. x -> (hi, lo), unless the caller prepared it with cgr_tc_split_features (batch preparation)
  if (need_x) {
    CgrRange prof("tc_split_x", st);
    cgr_note_launch("tc_split_x", st, 1);
    split_rows_kernel<<<(unsigned)cgr_ceil_div(N, 8), 256, 0, st>>>(g->x, fa, N, fa, x_hi, x_lo, w.kp_x, flag, FLAG_X);
    CGR_LAUNCH_CHECK();
  }
  // experiments only (results are wrong): CGR_DEBUG_SKIP=atom|edge|both leaves stages out to time the rest
  static const char* dbg_skip = getenv("CGR_DEBUG_SKIP");
  const bool skip_atom = dbg_skip && (!strcmp(dbg_skip, "atom") || !strcmp(dbg_skip, "both"));
  const bool skip_edge = dbg_skip && (!strcmp(dbg_skip, "edge") || !strcmp(dbg_skip, "both"));
  // 2. per-atom projections [P' | Q'] = x [W_x ; W_ox]^T + [b_i | b_o]   (GNN.py:86 and :106-107, x part)
  if (!skip_atom) {
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    if ((rc = make_map(&prm.tmA_hi, x_hi, N, fa, w.kp_x, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, x_lo, N, fa, w.kp_x, TM))) return rc;
    // throughput mode: wide slices (operand bytes streamed per batch: 42 MB instead of 67 MB at cfg-2)
    static const bool ap_wide = getenv("CGR_AP_NARROW") == nullptr;
    int bn = (p->tc_throughput != 0 && ap_wide) ? BN_LARGE : choose_bn(cgr_ceil_div(N, TM), 2 * H);
    const int pbn = proj_bn(bn, 2 * H);          // persistent kernel (wide-slice regime): its own slice width
    if (pbn) bn = pbn;
    if ((rc = make_map(&prm.tmB_hi, w_hi(0), 2 * H, fa, wl.ld[0], bn))) return rc;
    if ((rc = make_map(&prm.tmB_lo, w_lo(0), 2 * H, fa, wl.ld[0], bn))) return rc;
    prm.num_k = (int)cgr_ceil_div(fa, BK);
    prm.k_total = fa;
    prm.n_total = 2 * H;
    prm.m_rows = (int)N;
    prm.unscale = unscale + 0;
    prm.bias = bias_cat;
    prm.out_f32 = PQ;
    prm.ldc = 2 * H;
    prm.fast = fast;
    prm.overflow = flag;             // first kernel of the forward: clears the per-forward overflow bit
    if (pbn) {
      if ((rc = proj_multicast(&prm, w_hi(0), w_lo(0), 2 * H, fa, wl.ld[0], pbn))) return rc;
      rc = launch_proj(prm, pbn, (int)cgr_ceil_div(N, TM), st);
    } else {
      if (bn == BN_LARGE && (rc = ap_multicast(&prm, w_hi(0), w_lo(0), 2 * H, fa, wl.ld[0], bn))) return rc;
      rc = launch_gemm<EPI_PLAIN>(prm, bn, (int)cgr_ceil_div(N, TM), true, "tc_atom_proj", false,
                                  two_per_sm(cgr_ceil_div(N, TM) * cgr_ceil_div(2 * H, bn)), st);
    }
    if (rc) return rc;
  }
  const bool relu = p->act == CGR_ACT_RELU;
  // 4+5 fused: every bond layer and the readout of a tile group in ONE cluster launch (inference; the training forward
  // keeps one operand pair per layer for the backward and stays on the per-layer kernels)
  static const bool use_fused = getenv("CGR_NO_FUSED_FWD") == nullptr;
  static const bool use_fused_init = getenv("CGR_NO_FUSED_INIT") == nullptr;
  FwdChoice fc;
  bool any_dropout = false;                // dropout masks stay with the per-layer kernels (Philox stream per layer)
  for (int l = 0; training && p->host_dropout_p && l < d; ++l) any_dropout |= p->host_dropout_p[l] > 0.f;
  static const bool fused_train = getenv("CGR_NO_FUSED_TRAIN_FWD") == nullptr;
  const bool fused = use_fused && (!blob || fused_train) && !any_dropout && choose_fwd(T, H, p->tc_throughput != 0, &fc);
  // the fused kernel also computes h0 (edge initialisation) when W_e^T's slice and the tiles' bond features fit its
  // staging ring
  const bool fused_init = fused && use_fused_init && fb > 0 &&
                          (size_t)fb * (fc.bn + fc.tpc * TM) * sizeof(float) <= (size_t)tcf::ring_bytes(fc.bn);
  // 3. edge initialisation on tile-packed rows
  if (!skip_edge && !fused_init) {
    CgrRange prof("tc_edge_init", st);
    cgr_note_launch("tc_edge_init", st, 1);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(TM / EI_ROWS, (unsigned)T);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = (size_t)((fb > 0 ? fb : 1) * H + EI_ROWS * (fb > 0 ? fb : 1)) * sizeof(float);
    static bool ei_attr = false;
    if (!ei_attr) {
      CGR_CUDA(cudaFuncSetAttribute(tc_edge_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      ei_attr = true;
    }
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    CGR_CUDA(cudaLaunchKernelEx(&cfg, tc_edge_init_kernel, (const float*)PQ, (int64_t)(2 * H), g->edge_attr, g->src,
                                (const float*)(wbuf + wl.off_wet), g->tile_info, fb, H, (int)p->act, h0, hbuf_hi(0),
                                hbuf_lo(0), (int64_t)w.kp_h, flag));
  }
  if (fused) {
    tcf::FwdParams prm;
    memset(&prm, 0, sizeof(prm));
    if ((rc = fwd_fill_shared(&prm, p, wbuf, wl, fc, w.kp_h, fused_init, fast))) return rc;
    if ((rc = fwd_fill_batch(&prm.bt[0], g, ws, w, H, out))) return rc;
    if (blob && (rc = fwd_fill_training(&prm.bt[0], p, blob, SL, H))) return rc;
    prm.n_batches = 1;
    const int n_groups = (int)cgr_ceil_div(T, fc.tpc);
    return launch_fwd(prm, fc, relu, n_groups, use_pdl, st);
  }
  // 4. message passing layers: one fused kernel each
  const int bn_h = choose_bn(T, H);
  for (int l = 0; l < d; ++l) {
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    if ((rc = make_map(&prm.tmA_hi, hbuf_hi(l), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, hbuf_lo(l), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmB_hi, w_hi(1 + l), H, H, wl.ld[1 + l], bn_h))) return rc;
    if ((rc = make_map(&prm.tmB_lo, w_lo(1 + l), H, H, wl.ld[1 + l], bn_h))) return rc;
    prm.num_k = (int)cgr_ceil_div(H, BK);
    prm.k_total = H;
    prm.n_total = H;
    prm.unscale = unscale + 1 + l;
    prm.bias = p->b_conv[l];
    prm.tile_info = g->tile_info;
    prm.in_ptr = g->in_ptr; prm.in_idx = g->in_idx; prm.src = g->src; prm.atom_ptr = g->atom_ptr;
    prm.skip = p->use_skip ? p->skip[l] : nullptr;
    if ((rc = make_map_f32(&prm.tmR, h0, w.rows_pad, H, H, chunk_cols(bn_h), TM))) return rc;
    prm.r_col0 = 0;
    prm.act = p->act;
    prm.dropout_p = (training && p->host_dropout_p) ? p->host_dropout_p[l] : 0.f;
    prm.seed = seed; prm.layer = (uint32_t)l;
    prm.o_hi = hbuf_hi(l + 1); prm.o_lo = hbuf_lo(l + 1); prm.ldo = w.kp_h;
    prm.overflow = flag;
    prm.fast = fast;
    prm.dbg = g_tc_dbg;
    rc = launch_gemm<EPI_BOND>(prm, bn_h, (int)T, relu, "bond_layer", use_pdl, two_per_sm(T * cgr_ceil_div(H, bn_h)), st);
    if (rc) return rc;
  }
  // 5. readout + pooling + FFN
  {
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    if ((rc = make_map(&prm.tmA_hi, hbuf_hi(d), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, hbuf_lo(d), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmB_hi, w_hi(d + 1), H, H, wl.ld[d + 1], bn_h))) return rc;
    if ((rc = make_map(&prm.tmB_lo, w_lo(d + 1), H, H, wl.ld[d + 1], bn_h))) return rc;
    prm.num_k = (int)cgr_ceil_div(H, BK);
    prm.k_total = H;
    prm.n_total = H;
    prm.unscale = unscale + d + 1;
    prm.tile_info = g->tile_info;
    prm.in_ptr = g->in_ptr; prm.in_idx = g->in_idx; prm.src = g->src; prm.atom_ptr = g->atom_ptr;
    prm.act = p->act;
    if ((rc = make_map_f32(&prm.tmR, PQ, N, 2 * H, 2 * H, chunk_cols(bn_h), TM))) return rc;
    prm.r_col0 = H;
    prm.w_ffn = p->w_ffn;
    prm.b_ffn = p->b_ffn;
    prm.out = out;
    prm.tile_counter = tile_counter;
    prm.partial_out = partial;
    prm.n_rxn = B;
    prm.overflow = flag;
    prm.fast = fast;
    prm.hv_out = blob ? (float*)(blob + SL.off_hv) : nullptr;
    rc = launch_gemm<EPI_READOUT>(prm, bn_h, (int)T, relu, "tc_readout", use_pdl, two_per_sm(T * cgr_ceil_div(H, bn_h)), st);
    if (rc) return rc;
  }
  return CGR_OK;
}


// ------------------------------------------------------------------------------------------------
// Group forward: SEVERAL independent batches (a screening job is a stream of them) in TWO launches -- one atom
// projection over every batch's atom tiles, one fused cluster kernel over every batch's tile groups -- instead of
// two launches per batch pipelined over streams.  One grid packs the SMs wave after wave (a 64-reaction forward
// alone occupies 22 of the 148); per-batch operands, index arrays and outputs travel as kernel parameters
// (FwdBatch / GemmBatch tables in the constant bank), the weights are shared.  Same kernels, same arithmetic: the
// energies are bit-identical to per-batch forwards in the throughput configuration.
// ------------------------------------------------------------------------------------------------
namespace {
struct GroupWs { size_t off_w, off_g[tcf::MAX_GROUP], total; };
GroupWs group_ws(const cgr_params_t* p, const cgr_graph_t* gs, int n) {
  GroupWs G;
  size_t off = 0;
  G.off_w = off;
  off += cgr_align_up(p->tc_weights ? 0 : tc_weights_bytes(p), 1024);
  for (int i = 0; i < n; ++i) {
    G.off_g[i] = off;
    off += cgr_align_up(tc_ws(p, &gs[i], false).total, 1024);
  }
  G.total = off + 1024;
  return G;
}
}  // namespace

size_t tc_forward_group_workspace(const cgr_params_t* p, const cgr_graph_t* gs, int n) {
  if (n <= 0 || n > tcf::MAX_GROUP) return 0;
  for (int i = 0; i < n; ++i)
    if (!gs[i].tile_info || gs[i].n_tiles <= 0) return 0;
  return group_ws(p, gs, n).total;
}

int tc_gnn_forward_group(const cgr_params_t* p, const cgr_graph_t* gs, int n, float* const* outs, void* workspace,
                         size_t workspace_bytes, cudaStream_t st) {
  CGR_CHECK_ARG(n >= 1 && n <= tcf::MAX_GROUP, "group forward takes 1..%d batches per call", tcf::MAX_GROUP);
  CGR_CHECK_ARG(p->hidden % 4 == 0, "tcgen05 engine needs a hidden size that is a multiple of 4");
  CGR_CHECK_ARG(p->depth + 3 <= MAX_SEG, "tcgen05 engine supports depth <= %d", MAX_SEG - 3);
  CGR_CHECK_ARG(p->fb <= 32, "tcgen05 engine supports at most 32 bond features");
  const int H = p->hidden, fa = p->fa, fb = p->fb;
  int64_t t_sum = 0, m_tiles = 0;
  for (int i = 0; i < n; ++i) {
    CGR_CHECK_ARG(gs[i].tile_info && gs[i].n_tiles > 0 && gs[i].tc_status && outs[i],
                  "group forward: batch %d needs a tile plan, tc_status and an output", i);
    t_sum += gs[i].n_tiles;
    m_tiles += cgr_ceil_div(gs[i].n_atoms, TM);
  }
  CGR_CHECK_ARG(t_sum <= 60000 && m_tiles <= 60000, "group forward: too many row tiles for one launch");
  const GroupWs G = group_ws(p, gs, n);
  CGR_CHECK_ARG(workspace && workspace_bytes >= G.total, "tc_gnn_forward_group: workspace too small");
  char* base = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  int rc;
  char* wbuf = p->tc_weights ? (char*)p->tc_weights : base + G.off_w;
  if (!p->tc_weights && (rc = tc_prepare_weights(p, wbuf, tc_weights_bytes(p), st))) return rc;
  const WLayout wl = wlayout(p);
  const int fast = p->tc_fast ? 1 : 0;
  FwdChoice fc;
  CGR_CHECK_ARG(choose_fwd(2, H, true, &fc), "group forward: hidden size %d needs more than 8 column slices", H);
  const bool fused_init = fb > 0 && getenv("CGR_NO_FUSED_INIT") == nullptr &&
                          (size_t)fb * (fc.bn + fc.tpc * TM) * sizeof(float) <= (size_t)tcf::ring_bytes(fc.bn);
  static const bool use_pdl = getenv("CGR_NO_PDL") == nullptr;

  // 1 + 2. x -> (hi, lo) where the caller did not prepare it; ONE atom-projection launch over every batch's tiles
  TcGemmParams ap;
  memset(&ap, 0, sizeof(ap));
  tcf::FwdParams prm;
  memset(&prm, 0, sizeof(prm));
  TcWs wi[tcf::MAX_GROUP];
  int tile0 = 0, group0 = 0;
  for (int i = 0; i < n; ++i) {
    const cgr_graph_t* g = &gs[i];
    wi[i] = tc_ws(p, g, false);
    const TcWs& w = wi[i];
    char* ws = base + G.off_g[i];
    const bool need_x = !(g->x_hi && g->x_lo);
    __half* x_hi = need_x ? (__half*)(ws + w.off_xhi) : (__half*)g->x_hi;
    __half* x_lo = need_x ? (__half*)(ws + w.off_xlo) : (__half*)g->x_lo;
    if (need_x) {
      CgrRange prof("tc_split_x", st);
      cgr_note_launch("tc_split_x", st, 1);
      split_rows_kernel<<<(unsigned)cgr_ceil_div(g->n_atoms, 8), 256, 0, st>>>(g->x, fa, g->n_atoms, fa, x_hi, x_lo, w.kp_x,
                                                                              g->tc_status, FLAG_X);
      CGR_LAUNCH_CHECK();
    }
    GemmBatch& gb = ap.gb[i];
    if ((rc = make_map(&gb.tmA_hi, x_hi, g->n_atoms, fa, w.kp_x, TM))) return rc;
    if ((rc = make_map(&gb.tmA_lo, x_lo, g->n_atoms, fa, w.kp_x, TM))) return rc;
    gb.out_f32 = (float*)(ws + w.off_pq);
    gb.overflow = g->tc_status;
    gb.m_rows = (int)g->n_atoms;
    gb.tile0 = tile0;
    tile0 += (int)cgr_ceil_div(g->n_atoms, TM);
    if ((rc = fwd_fill_batch(&prm.bt[i], g, ws, w, H, outs[i]))) return rc;
    prm.bt[i].group0 = group0;
    group0 += (int)cgr_ceil_div(g->n_tiles, fc.tpc);
  }
  {
    const int pbn = proj_bn(BN_LARGE, 2 * H);
    const int bn = pbn ? pbn : BN_LARGE;
    if ((rc = make_map(&ap.tmB_hi, (const __half*)(wbuf + wl.off_hi[0]), 2 * H, fa, wl.ld[0], bn))) return rc;
    if ((rc = make_map(&ap.tmB_lo, (const __half*)(wbuf + wl.off_lo[0]), 2 * H, fa, wl.ld[0], bn))) return rc;
    ap.tmA_hi = ap.gb[0].tmA_hi; ap.tmA_lo = ap.gb[0].tmA_lo;
    ap.num_k = (int)cgr_ceil_div(fa, BK);
    ap.k_total = fa;
    ap.n_total = 2 * H;
    ap.unscale = (const float*)(wbuf + wl.off_unscale);
    ap.bias = (const float*)(wbuf + wl.off_bias);
    ap.ldc = 2 * H;
    ap.fast = fast;
    ap.n_batches = n;
    if (pbn) {
      if ((rc = proj_multicast(&ap, (const __half*)(wbuf + wl.off_hi[0]), (const __half*)(wbuf + wl.off_lo[0]), 2 * H, fa,
                               wl.ld[0], pbn))) return rc;
      rc = launch_proj(ap, pbn, tile0, st);
    } else {
      if ((rc = ap_multicast(&ap, (const __half*)(wbuf + wl.off_hi[0]), (const __half*)(wbuf + wl.off_lo[0]), 2 * H, fa,
                             wl.ld[0], bn))) return rc;
      rc = launch_gemm<EPI_PLAIN>(ap, bn, tile0, true, "tc_atom_proj", false, false, st);
    }
    if (rc) return rc;
  }
  // 3. edge initialisation: inside the fused kernel, or (bond features too wide for its staging ring) per batch
  if (!fused_init) {
    for (int i = 0; i < n; ++i) {
      const cgr_graph_t* g = &gs[i];
      const TcWs& w = wi[i];
      char* ws = base + G.off_g[i];
      CgrRange prof("tc_edge_init", st);
      cgr_note_launch("tc_edge_init", st, 1);
      static bool ei_attr = false;
      if (!ei_attr) {
        CGR_CUDA(cudaFuncSetAttribute(tc_edge_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        ei_attr = true;
      }
      const size_t smem = (size_t)((fb > 0 ? fb : 1) * H + EI_ROWS * (fb > 0 ? fb : 1)) * sizeof(float);
      tc_edge_init_kernel<<<dim3(TM / EI_ROWS, (unsigned)g->n_tiles), 256, smem, st>>>(
          (const float*)(ws + w.off_pq), (int64_t)(2 * H), g->edge_attr, g->src, (const float*)(wbuf + wl.off_wet),
          g->tile_info, fb, H, (int)p->act, (float*)(ws + w.off_h0), (__half*)(ws + w.off_hhi[0]),
          (__half*)(ws + w.off_hlo[0]), (int64_t)w.kp_h, g->tc_status);
      CGR_LAUNCH_CHECK();
    }
  }
  // 4 + 5. every bond layer and the readout of every batch's tile groups in one cluster launch
  if ((rc = fwd_fill_shared(&prm, p, wbuf, wl, fc, wi[0].kp_h, fused_init, fast))) return rc;
  prm.n_batches = n;
  return launch_fwd(prm, fc, p->act == CGR_ACT_RELU, group0, use_pdl && fused_init, st);
}


// ------------------------------------------------------------------------------------------------
// Fused tile-local backward (training on the tcgen05 engine, ReLU networks on tileable batches).
//
// With the GEMM-first formulation  y_l = h_{l-1} W_l^T,  z_l = G y_l + b_l + skip_l h_0,  h_l = drop(relu(z_l))
// (G = directed-bond gather), the backward of a layer is  dz_l = dh_l . [h_l > 0] . keep_scale,
// dy_l = G^T dz_l,  dW_l = dy_l^T h_{l-1},  dh_{l-1} = dy_l W_l:  the same shape as the forward -- one GEMM whose
// epilogue masks, reduces and gathers inside the tile (EPI_BOND_BWD) -- so the only tensors saved are the layer
// outputs the forward already writes as FP16 (hi, lo) operands.  Gradient operands carry a power-of-two scale taken
// from the amax of the previous gradient tensor (max is order-independent: deterministic).
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int BWD_THREADS = 256;
constexpr int RB_PARTS = 8;          // blocks per tile in the readout backward

// gamax[0] = max|dout| * max|w_f| (bound on |dq|), gamax[1..] = 0;  db_f = sum_b dout[b]
__global__ void __launch_bounds__(BWD_THREADS) bwd_prep_kernel(const float* __restrict__ dout, int64_t B,
                                                               const float* __restrict__ w_ffn, int H,
                                                               unsigned int* __restrict__ gamax, int n_gamax,
                                                               float* __restrict__ db_ffn) {
  __shared__ float red[BWD_THREADS];
  __shared__ float red2[BWD_THREADS];
  float am = 0.f, sum = 0.f, wm = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += BWD_THREADS) {
    const float v = __ldg(dout + i);
    am = fmaxf(am, fabsf(v));
    sum += v;
  }
  for (int i = threadIdx.x; i < H; i += BWD_THREADS) wm = fmaxf(wm, fabsf(__ldg(w_ffn + i)));
  red[threadIdx.x] = sum;
  red2[threadIdx.x] = am;
  __syncthreads();
  for (int o = BWD_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[threadIdx.x] += red[threadIdx.x + o];
      red2[threadIdx.x] = fmaxf(red2[threadIdx.x], red2[threadIdx.x + o]);
    }
    __syncthreads();
  }
  const float amax_d = red2[0];
  if (threadIdx.x == 0) *db_ffn = red[0];
  __syncthreads();
  red2[threadIdx.x] = wm;
  __syncthreads();
  for (int o = BWD_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red2[threadIdx.x] = fmaxf(red2[threadIdx.x], red2[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) gamax[0] = __float_as_uint(amax_d * red2[0]);
  for (int i = 1 + threadIdx.x; i < n_gamax; i += BWD_THREADS) gamax[i] = 0u;
}

// Readout backward, RB_PARTS blocks per tile (rows interleaved over the blocks):  dzv[v] = dout[rxn(v)] . w_f . [hv[v] > 0]  (atoms of the tile),
// dq[k] = dzv[dst k] (bonds of the tile, tile-packed), both as scaled FP16 (hi, lo) operands; per-tile column sums
// of dzv (db_o) and of dout . hv (dw_f).
__global__ void __launch_bounds__(BWD_THREADS) readout_bwd_kernel(
    const float* __restrict__ dout, const float* __restrict__ hv, const float* __restrict__ w_ffn,
    const int32_t* __restrict__ tile_info, const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ dst, int H,
    int64_t kp_h, __half* __restrict__ dzv_hi, __half* __restrict__ dzv_lo, __half* __restrict__ dq_hi,
    __half* __restrict__ dq_lo, float* __restrict__ dbo_partial, float* __restrict__ dwf_partial,
    unsigned int* __restrict__ gamax, float* __restrict__ gunscale) {
  extern __shared__ __align__(16) float rb_smem[];          // [8][H] cross-warp reduction scratch
  __shared__ float dpl[TM];
  const int tile = blockIdx.y, part = blockIdx.x;
  const int ebase = __ldg(tile_info + tile * 8), ecount = __ldg(tile_info + tile * 8 + 1);
  const int abase = __ldg(tile_info + tile * 8 + 2), acount = __ldg(tile_info + tile * 8 + 3);
  const int rx0 = __ldg(tile_info + tile * 8 + 4), rxcount = __ldg(tile_info + tile * 8 + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ROW_STEP = RB_PARTS * (BWD_THREADS / 32);
  const int row_first = part + RB_PARTS * warp;              // this warp's rows: row_first, + ROW_STEP, ...
  for (int rx = threadIdx.x; rx < rxcount; rx += BWD_THREADS) {
    const float d = __ldg(dout + rx0 + rx);
    const int v0 = __ldg(atom_ptr + rx0 + rx) - abase, v1 = __ldg(atom_ptr + rx0 + rx + 1) - abase;
    for (int v = v0; v < v1; ++v) dpl[v] = d;
  }
  const float S0 = tcg::grad_scale(gamax[0]);
  if (tile == 0 && part == 0 && threadIdx.x == 0) gunscale[0] = 1.f / S0;
  __syncthreads();
  constexpr int NG = 8;                                      // column groups of 128: H <= 1024
  float4 csum[NG], wsum[NG];
#pragma unroll
  for (int q = 0; q < NG; ++q) csum[q] = wsum[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = row_first; v < acount; v += ROW_STEP) {
    const float d = dpl[v];
    const int64_t row = abase + v;
#pragma unroll
    for (int q = 0; q < NG; ++q) {
      const int n = q * 128 + 4 * lane;
      if (n < H) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(hv + row * H + n));
        const float4 w = __ldg(reinterpret_cast<const float4*>(w_ffn + n));
        float4 dz;
        dz.x = h.x > 0.f ? d * w.x : 0.f; dz.y = h.y > 0.f ? d * w.y : 0.f;
        dz.z = h.z > 0.f ? d * w.z : 0.f; dz.w = h.w > 0.f ? d * w.w : 0.f;
        tcg::store_split4(dz, S0, dzv_hi + row * kp_h + n, dzv_lo + row * kp_h + n);
        tcg::add4(csum[q], dz);
        wsum[q].x = fmaf(d, h.x, wsum[q].x); wsum[q].y = fmaf(d, h.y, wsum[q].y);
        wsum[q].z = fmaf(d, h.z, wsum[q].z); wsum[q].w = fmaf(d, h.w, wsum[q].w);
      }
    }
  }
  for (int pass = 0; pass < 2; ++pass) {                     // fixed warp order: deterministic
#pragma unroll
    for (int q = 0; q < NG; ++q) {
      const int n = q * 128 + 4 * lane;
      if (n < H) *reinterpret_cast<float4*>(rb_smem + warp * H + n) = pass == 0 ? csum[q] : wsum[q];
    }
    __syncthreads();
    float* outp = pass == 0 ? dbo_partial : dwf_partial;
    for (int n = threadIdx.x; n < H; n += BWD_THREADS) {
      float t = 0.f;
      for (int w = 0; w < BWD_THREADS / 32; ++w) t += rb_smem[w * H + n];
      outp[((int64_t)tile * RB_PARTS + part) * H + n] = t;
    }
    __syncthreads();
  }
  float vmax = 0.f;
  for (int j = row_first; j < ecount; j += ROW_STEP) {
    const int v = __ldg(dst + ebase + j) - abase;
    const float d = dpl[v];
    const int64_t row = abase + v, orow = (int64_t)tile * TM + j;
    for (int n = 4 * lane; n < H; n += 128) {
      const float4 h = __ldg(reinterpret_cast<const float4*>(hv + row * H + n));
      const float4 w = __ldg(reinterpret_cast<const float4*>(w_ffn + n));
      float4 dz;
      dz.x = h.x > 0.f ? d * w.x : 0.f; dz.y = h.y > 0.f ? d * w.y : 0.f;
      dz.z = h.z > 0.f ? d * w.z : 0.f; dz.w = h.w > 0.f ? d * w.w : 0.f;
      vmax = fmaxf(vmax, tcg::amax4(dz));
      tcg::store_split4(dz, S0, dq_hi + orow * kp_h + n, dq_lo + orow * kp_h + n);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if (lane == 0 && vmax > 0.f) atomicMax(gamax + 1, __float_as_uint(vmax));
}

// Sums the per-tile partials in tile order: bias gradients / dw_f  ([T][H] -> [H]) and skip-weight gradients.
struct BwdFinalizeArgs {
  const float* col_src[MAX_SEG + 2];
  float* col_dst[MAX_SEG + 2];
  int col_rows[MAX_SEG + 2];
  int n_col;
  const float* skip_src[MAX_SEG];
  float* skip_dst[MAX_SEG];
  int n_skip;
  int skip_cnt;                 // values per skip partial array
  int T, H;
};
__global__ void __launch_bounds__(256) bwd_finalize_kernel(const BwdFinalizeArgs a) {
  if ((int)blockIdx.y < a.n_col) {
    // 32 columns per block; 8 row lanes stride over the partial rows, then a fixed-order combine: deterministic
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cx;
    const float* src = a.col_src[blockIdx.y];
    const int rows = a.col_rows[blockIdx.y];
    float t = 0.f;
    if (n < a.H)
      for (int i = ry; i < rows; i += 8) t += __ldg(src + (int64_t)i * a.H + n);
    red[ry][cx] = t;
    __syncthreads();
    if (ry == 0 && n < a.H) {
      float sres = 0.f;
      for (int w = 0; w < 8; ++w) sres += red[w][cx];
      a.col_dst[blockIdx.y][n] = sres;
    }
  } else if (blockIdx.x == 0 && threadIdx.x < 32) {
    const int k = blockIdx.y - a.n_col;
    const float* src = a.skip_src[k];
    float t = 0.f;
    for (int i = threadIdx.x; i < a.skip_cnt; i += 32) t += __ldg(src + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) *a.skip_dst[k] = t;
  }
}

struct TcBwdWs {
  int64_t kp_h, kp_x, rows_pad;
  int n_slices;
  size_t off_ghi[MAX_SEG], off_glo[MAX_SEG], off_dzv_hi, off_dzv_lo, off_dp_hi, off_dp_lo, zero_bytes;
  size_t off_dh0, off_dz0, off_col, off_dwf, off_skip, off_gamax, off_gunscale, off_partial, off_xhi, off_xlo, off_w, total;
  size_t skip_cnt, partial_floats;
};
TcBwdWs tc_bwd_ws(const cgr_params_t* p, const cgr_graph_t* g) {
  TcBwdWs w;
  const int64_t H = p->hidden, N = g->n_atoms, E = g->n_bonds, T = g->n_tiles;
  const int d = p->depth;
  w.kp_h = round_up(H, BK);
  w.kp_x = round_up(p->fa, BK);
  w.rows_pad = T * TM;
  w.n_slices = (int)cgr_ceil_div(H, BN_SMALL);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += cgr_align_up(bytes, 1024); return o; };
  for (int i = 0; i <= d; ++i) {                       // G_0 = dq, G_i = dy_{d+1-i}
    w.off_ghi[i] = take((size_t)w.rows_pad * w.kp_h * sizeof(__half));
    w.off_glo[i] = take((size_t)w.rows_pad * w.kp_h * sizeof(__half));
  }
  w.zero_bytes = off;                                  // tile-packed operands: pad rows must be zero (K of the wgrads)
  w.off_dzv_hi = take((size_t)N * w.kp_h * sizeof(__half));
  w.off_dzv_lo = take((size_t)N * w.kp_h * sizeof(__half));
  w.off_dp_hi = take((size_t)N * w.kp_h * sizeof(__half));
  w.off_dp_lo = take((size_t)N * w.kp_h * sizeof(__half));
  w.off_dh0 = take((size_t)w.rows_pad * H * sizeof(float));
  w.off_dz0 = take((size_t)E * H * sizeof(float));
  w.off_col = take((size_t)(d + 1 + RB_PARTS) * T * H * sizeof(float));   // db_o [T*RB_PARTS][H], then d+1 arrays [T][H]
  w.off_dwf = take((size_t)T * RB_PARTS * H * sizeof(float));
  w.skip_cnt = (size_t)T * w.n_slices * 2;
  w.off_skip = take((size_t)d * w.skip_cnt * sizeof(float));
  w.off_gamax = take(MAX_SEG * 2 * sizeof(unsigned int));
  w.off_gunscale = take(MAX_SEG * 2 * sizeof(float));
  size_t pf = 0;
  auto upd = [&](int n_max, int64_t M_, int64_t N_, int64_t K_) {    // batched weight-gradient GEMMs (any batch <= n_max)
    for (int n_ = 1; n_ <= n_max && n_ <= tcg2::MAX_BATCH; ++n_) {
      const int sk = tc_splitk_batched(n_, M_, N_, K_);
      const size_t f = sk > 1 ? (size_t)n_ * sk * M_ * N_ : 0;
      if (f > pf) pf = f;
    }
  };
  upd(d + 1, H, H, w.rows_pad);
  upd(2, H, p->fa, N);
  const int ssk = simt_splitk_choose(H, p->fb > 0 ? p->fb : 1, E);
  if (ssk > 1 && (size_t)ssk * H * p->fb > pf) pf = (size_t)ssk * H * p->fb;
  w.partial_floats = pf;
  w.off_partial = take(pf * sizeof(float));
  const bool need_x = !(g->x_hi && g->x_lo);
  w.off_xhi = take(need_x ? (size_t)N * w.kp_x * sizeof(__half) : 0);
  w.off_xlo = take(need_x ? (size_t)N * w.kp_x * sizeof(__half) : 0);
  w.off_w = 0;
  w.total = off + 1024;
  return w;
}

}  // namespace

size_t tc_backward_workspace(const cgr_params_t* p, const cgr_graph_t* g) {
  if (!tc_fused_training_ok(p, g)) return 0;
  return tc_bwd_ws(p, g).total;
}

namespace {
// The weight-gradient GEMMs only consume what the backward chain produces: they run on a side stream, each as soon
// as its operand exists, next to the (latency-bound, ~100-CTA) chain kernels instead of after them.  Fork and join
// are event dependencies, so a stream capture records them as parallel branches of the step's graph.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[MAX_SEG + 4] = {};
};
SideStream* bwd_side_stream() {
  // measured (cfg-3, B = 64, one CUDA graph per step): 0.342 ms forked against 0.315-0.329 ms in line -- the step is a
  // chain of latency-bound kernels, the GEMMs compete with it for SMs and add split-K reductions; off unless asked for
  static const bool on = getenv("CGR_BWD_FORK") != nullptr;
  if (!on) return nullptr;
  thread_local SideStream per_dev[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& S = per_dev[dev];
  if (!S.s) {
    if (cudaStreamCreateWithFlags(&S.s, cudaStreamNonBlocking) != cudaSuccess) { S.s = nullptr; return nullptr; }
    for (auto& e : S.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &S;
}
}  // namespace

int tc_gnn_backward(const cgr_params_t* p, const cgr_graph_t* g, const cgr_saved_t* saved, const float* dout,
                    const cgr_grads_t* grads, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CGR_CHECK_ARG(tc_fused_training_ok(p, g) && saved && saved->tc_blob, "tcgen05 fused backward: unsupported configuration");
  CGR_CHECK_ARG(g->n_tiles <= 65535, "tcgen05 engine: too many row tiles for one launch: split the batch");
  const TcSavedLayout SL = tc_saved_layout(p, g);
  const TcBwdWs w = tc_bwd_ws(p, g);
  CGR_CHECK_ARG(saved->tc_blob_bytes >= SL.total, "tcgen05 fused backward: tc_blob too small");
  CGR_CHECK_ARG(workspace && workspace_bytes >= w.total, "tcgen05 fused backward: workspace too small");
  CGR_CHECK_ARG(g->tc_status && g->dst, "tcgen05 fused backward: tc_status / dst missing");
  char* blob = (char*)(((uintptr_t)saved->tc_blob + 1023) & ~(uintptr_t)1023);
  char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const int H = p->hidden, fa = p->fa, fb = p->fb, d = p->depth;
  const int64_t N = g->n_atoms, E = g->n_bonds, T = g->n_tiles, B = g->n_rxn;
  int rc;
  // weights: the caller's prepared buffer, else the copy the training forward left in the blob
  char* wbuf = p->tc_weights ? (char*)p->tc_weights : blob + SL.off_w;
  const WLayout wl = wlayout(p);
  const float* unscale = (const float*)(wbuf + wl.off_unscale);
  int* flag = g->tc_status;
  const __half* x_hi = (const __half*)g->x_hi;
  const __half* x_lo = (const __half*)g->x_lo;
  if (!(x_hi && x_lo)) {
    __half* xh = (__half*)(ws + w.off_xhi);
    __half* xl = (__half*)(ws + w.off_xlo);
    if ((rc = tc_split_features(g->x, N, fa, xh, xl, flag, st))) return rc;
    x_hi = xh; x_lo = xl;
  }
  auto h_hi = [&](int l) { return (const __half*)(blob + SL.off_hhi[l]); };
  auto h_lo = [&](int l) { return (const __half*)(blob + SL.off_hlo[l]); };
  const float* h0 = (const float*)(blob + SL.off_h0);
  const float* hv = (const float*)(blob + SL.off_hv);
  auto g_hi = [&](int i) { return (__half*)(ws + w.off_ghi[i]); };
  auto g_lo = [&](int i) { return (__half*)(ws + w.off_glo[i]); };
  __half* dzv_hi = (__half*)(ws + w.off_dzv_hi);
  __half* dzv_lo = (__half*)(ws + w.off_dzv_lo);
  __half* dp_hi = (__half*)(ws + w.off_dp_hi);
  __half* dp_lo = (__half*)(ws + w.off_dp_lo);
  float* dh0 = (float*)(ws + w.off_dh0);
  float* dz0 = (float*)(ws + w.off_dz0);
  float* col = (float*)(ws + w.off_col);               // [d+2][T][H]: 0 = db_o, 1..d = db_conv[d-1..0] in launch order, d+1 = db_i
  float* dwf = (float*)(ws + w.off_dwf);
  float* skp = (float*)(ws + w.off_skip);
  unsigned int* gamax = (unsigned int*)(ws + w.off_gamax);
  float* gunscale = (float*)(ws + w.off_gunscale);
  float* partial = (float*)(ws + w.off_partial);
  static const bool use_pdl = getenv("CGR_NO_PDL") == nullptr;
  auto two_per_sm = [&](int64_t ctas) { return p->tc_throughput != 0 || ctas > 148; };

  CGR_CUDA(cudaMemsetAsync(ws, 0, w.zero_bytes, st));
  {
    CgrRange prof("bwd_readout", st);
    cgr_note_launch("bwd_readout", st, 2);
    bwd_prep_kernel<<<1, BWD_THREADS, 0, st>>>(dout, B, p->w_ffn, H, gamax, d + 4, grads->b_ffn);
    static bool attr = false;
    if (!attr) {
      CGR_CUDA(cudaFuncSetAttribute(readout_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 4));
      attr = true;
    }
    readout_bwd_kernel<<<dim3(RB_PARTS, (unsigned)T), BWD_THREADS, (size_t)8 * H * sizeof(float), st>>>(
        dout, hv, p->w_ffn, g->tile_info, g->atom_ptr, g->dst, H, w.kp_h, dzv_hi, dzv_lo, g_hi(0), g_lo(0), col, dwf,
        gamax, gunscale);
    CGR_LAUNCH_CHECK();
  }
  // weight gradients: reductions over bonds / atoms (K = rows) on tensor cores, deterministic split-K.  With the side
  // stream every GEMM is issued as soon as its gradient operand exists; without it they follow the chain on `st`.
  SideStream* side = bwd_side_stream();
  cudaStream_t ws_st = side ? side->s : st;
  int n_ev = 0;
  auto fork = [&]() -> int {                        // the side stream's next launches wait for everything issued on st so far
    if (!side) return CGR_OK;
    CGR_CUDA(cudaEventRecord(side->ev[n_ev], st));
    CGR_CUDA(cudaStreamWaitEvent(side->s, side->ev[n_ev], 0));
    ++n_ev;
    return CGR_OK;
  };
  auto wgrad = [&](const TcOperand& A_, const TcOperand& B_, float* C_, int64_t ldc_, int64_t M_, int64_t N_, int64_t K_,
                   const char* tag) {
    // forked: few CTAs per GEMM (it shares the machine with the chain kernel that runs next to it and must not take
    // the SMs that kernel needs); otherwise the machine-filling split
    int sk = tc_splitk_batched(1, M_, N_, K_);
    if (side) {
      static const int fork_ctas = getenv("CGR_BWD_FORK_CTAS") ? atoi(getenv("CGR_BWD_FORK_CTAS")) : 48;
      const int64_t base = cgr_ceil_div(M_, tcg2::TM) * cgr_ceil_div(N_, tcg2::TN);
      const int cap = (int)(fork_ctas / base > 1 ? fork_ctas / base : 1);
      if (sk > cap) sk = cap;
    }
    return tc_train_gemm_batched_mn(&A_, &B_, &C_, &ldc_, 1, M_, N_, K_, sk, partial, tag, ws_st);
  };
  if (side) {
    // dq and dzv exist: dW_os = dq^T h_d, dW_ox = dzv^T x
    if ((rc = fork())) return rc;
    if ((rc = wgrad(TcOperand{g_hi(0), g_lo(0), w.kp_h, gunscale + 0, true}, TcOperand{h_hi(d), h_lo(d), w.kp_h, nullptr, true},
                    grads->w_e2n + fa, fa + H, H, H, w.rows_pad, "wgrad_bond"))) return rc;
    if ((rc = wgrad(TcOperand{dzv_hi, dzv_lo, w.kp_h, gunscale + 0, true}, TcOperand{x_hi, x_lo, w.kp_x, nullptr, true},
                    grads->w_e2n, fa + H, H, fa, N, "wgrad_atoms"))) return rc;
  }
  const int bn_h = choose_bn(T, H);
  // bond layers, last to first: kernel i turns G_i (dq or dy_{l+1}) into dy_l = G_{i+1}, l = d - i
  for (int i = 0; i < d; ++i) {
    const int l = d - i;                                 // 1-based layer whose output gradient this GEMM produces
    const int wm = i == 0 ? d + 1 : l + 1;               // W_os for the readout, else W_{l+1} (matrix index = 1-based layer)
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    if ((rc = make_map(&prm.tmA_hi, g_hi(i), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, g_lo(i), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmB_hi, (const __half*)(wbuf + wl.off_hiT[wm]), H, H, wl.ld[wm], bn_h))) return rc;
    if ((rc = make_map(&prm.tmB_lo, (const __half*)(wbuf + wl.off_loT[wm]), H, H, wl.ld[wm], bn_h))) return rc;
    prm.num_k = (int)cgr_ceil_div(H, BK);
    prm.k_total = H;
    prm.n_total = H;
    prm.unscale = unscale + wm;
    prm.tile_info = g->tile_info;
    prm.in_ptr = g->in_ptr; prm.in_idx = g->in_idx; prm.src = g->src; prm.atom_ptr = g->atom_ptr;
    prm.skip = p->use_skip ? p->skip[l - 1] : nullptr;
    if ((rc = make_map_f32(&prm.tmR, h0, w.rows_pad, H, H, chunk_cols(bn_h), TM))) return rc;
    prm.act = p->act;
    prm.mask_hi = h_hi(l); prm.ld_mask = w.kp_h;
    const float pd = p->host_dropout_p ? p->host_dropout_p[l - 1] : 0.f;
    prm.keep_scale = pd > 0.f ? 1.f / (1.f - pd) : 1.f;
    prm.dh0_acc = dh0; prm.dh0_first = i == 0 ? 1 : 0;
    prm.colsum_partial = col + (size_t)(RB_PARTS + i) * T * H;
    prm.dskip_partial = p->use_skip ? skp + (size_t)i * w.skip_cnt : nullptr;
    prm.gamax_in = gamax + i; prm.gamax_out = gamax + i + 1; prm.gamax_track = gamax + i + 2;
    prm.gunscale_out = gunscale + i + 1;
    prm.o_hi = g_hi(i + 1); prm.o_lo = g_lo(i + 1); prm.ldo = w.kp_h;
    prm.overflow = flag;
    rc = launch_gemm<EPI_BOND_BWD>(prm, bn_h, (int)T, true, "bwd_bond_layer", use_pdl && i > 0,
                                   two_per_sm(T * cgr_ceil_div(H, bn_h)), st);
    if (rc) return rc;
    if (side) {
      // dy_l exists: dW_l = dy_l^T h_{l-1}
      if ((rc = fork())) return rc;
      if ((rc = wgrad(TcOperand{g_hi(i + 1), g_lo(i + 1), w.kp_h, gunscale + i + 1, true},
                      TcOperand{h_hi(l - 1), h_lo(l - 1), w.kp_h, nullptr, true}, grads->w_conv[l - 1], H, H, H, w.rows_pad,
                      "wgrad_bond"))) return rc;
    }
  }
  // edge initialisation: dh_0 = dy_1 W_1 + dh0_acc, dz_0 = dh_0 . [h_0 > 0], dP[v] = sum of dz_0 over bonds leaving v
  {
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    if ((rc = make_map(&prm.tmA_hi, g_hi(d), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, g_lo(d), w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmB_hi, (const __half*)(wbuf + wl.off_hiT[1]), H, H, wl.ld[1], bn_h))) return rc;
    if ((rc = make_map(&prm.tmB_lo, (const __half*)(wbuf + wl.off_loT[1]), H, H, wl.ld[1], bn_h))) return rc;
    prm.num_k = (int)cgr_ceil_div(H, BK);
    prm.k_total = H;
    prm.n_total = H;
    prm.unscale = unscale + 1;
    prm.tile_info = g->tile_info;
    prm.in_ptr = g->in_ptr; prm.in_idx = g->in_idx; prm.src = g->src; prm.atom_ptr = g->atom_ptr;
    if ((rc = make_map_f32(&prm.tmR, dh0, w.rows_pad, H, H, chunk_cols(bn_h), TM))) return rc;
    prm.act = p->act;
    prm.mask_hi = h_hi(0); prm.ld_mask = w.kp_h;
    prm.keep_scale = 1.f;
    prm.colsum_partial = col + (size_t)(RB_PARTS + d) * T * H;
    prm.gamax_in = gamax + d; prm.gamax_out = gamax + d + 1; prm.gamax_track = gamax + d + 2;
    prm.gunscale_out = gunscale + d + 1;
    prm.o_hi = dp_hi; prm.o_lo = dp_lo; prm.ldo = w.kp_h;
    prm.dz0_out = dz0;
    prm.overflow = flag;
    rc = launch_gemm<EPI_INIT_BWD>(prm, bn_h, (int)T, true, "bwd_edge_init", use_pdl, two_per_sm(T * cgr_ceil_div(H, bn_h)), st);
    if (rc) return rc;
  }
  if (side) {
    // dP and dz_0 exist: dW_x = dP^T x here, dW_e below (ws_st)
    if ((rc = fork())) return rc;
    if ((rc = wgrad(TcOperand{dp_hi, dp_lo, w.kp_h, gunscale + d + 1, true}, TcOperand{x_hi, x_lo, w.kp_x, nullptr, true},
                    grads->w_init, fa + fb, H, fa, N, "wgrad_atoms"))) return rc;
  } else {
    // without the side stream all GEMMs of one shape share a launch
    TcOperand A[tcg2::MAX_BATCH], Bo[tcg2::MAX_BATCH];
    float* C[tcg2::MAX_BATCH];
    int64_t ldc[tcg2::MAX_BATCH];
    int n = 0;
    auto flush = [&](int64_t M_, int64_t N_, int64_t K_, const char* tag) {
      if (n == 0) return (int)CGR_OK;
      const int r = tc_train_gemm_batched_mn(A, Bo, C, ldc, n, M_, N_, K_, tc_splitk_batched(n, M_, N_, K_), partial, tag, st);
      n = 0;
      return r;
    };
    // dW_os = dq^T h_d ;  dW_l = dy_l^T h_{l-1}
    A[n] = TcOperand{g_hi(0), g_lo(0), w.kp_h, gunscale + 0, true};
    Bo[n] = TcOperand{h_hi(d), h_lo(d), w.kp_h, nullptr, true};
    C[n] = grads->w_e2n + fa; ldc[n] = fa + H; ++n;
    for (int l = 1; l <= d; ++l) {
      if (n == tcg2::MAX_BATCH && (rc = flush(H, H, w.rows_pad, "wgrad_bond"))) return rc;
      const int i = d + 1 - l;
      A[n] = TcOperand{g_hi(i), g_lo(i), w.kp_h, gunscale + i, true};
      Bo[n] = TcOperand{h_hi(l - 1), h_lo(l - 1), w.kp_h, nullptr, true};
      C[n] = grads->w_conv[l - 1]; ldc[n] = H; ++n;
    }
    if ((rc = flush(H, H, w.rows_pad, "wgrad_bond"))) return rc;
    // dW_ox = dzv^T x ;  dW_x = dP^T x
    A[n] = TcOperand{dzv_hi, dzv_lo, w.kp_h, gunscale + 0, true};
    Bo[n] = TcOperand{x_hi, x_lo, w.kp_x, nullptr, true};
    C[n] = grads->w_e2n; ldc[n] = fa + H; ++n;
    A[n] = TcOperand{dp_hi, dp_lo, w.kp_h, gunscale + d + 1, true};
    Bo[n] = TcOperand{x_hi, x_lo, w.kp_x, nullptr, true};
    C[n] = grads->w_init; ldc[n] = fa + fb; ++n;
    if ((rc = flush(H, fa, N, "wgrad_atoms"))) return rc;
  }
  if (fb > 0) {        // [H x fb] with fb = 14: too narrow for a tensor-core tile, stays on the fp32 kernel
    GemmEpilogue e;
    e.tag = "wgrad_edge_attr";
    rc = simt_gemm(dz0, H, false, g->edge_attr, fb, false, grads->w_init + fa, fa + fb, H, fb, E, e,
                   simt_splitk_choose(H, fb, E), partial, ws_st);
    if (rc) return rc;
  }
  {
    BwdFinalizeArgs a;
    memset(&a, 0, sizeof(a));
    a.T = (int)T; a.H = H;
    int nc = 0;
    a.col_src[nc] = col; a.col_rows[nc] = (int)T * RB_PARTS; a.col_dst[nc++] = grads->b_e2n;
    for (int i = 0; i < d; ++i) {
      a.col_src[nc] = col + (size_t)(RB_PARTS + i) * T * H; a.col_rows[nc] = (int)T; a.col_dst[nc++] = grads->b_conv[d - 1 - i];
    }
    a.col_src[nc] = col + (size_t)(RB_PARTS + d) * T * H; a.col_rows[nc] = (int)T; a.col_dst[nc++] = grads->b_init;
    a.col_src[nc] = dwf; a.col_rows[nc] = (int)T * RB_PARTS; a.col_dst[nc++] = grads->w_ffn;
    a.n_col = nc;
    if (p->use_skip) {
      for (int i = 0; i < d; ++i) { a.skip_src[i] = skp + (size_t)i * w.skip_cnt; a.skip_dst[i] = grads->skip[d - 1 - i]; }
      a.n_skip = d;
      // each launch wrote [T][n_slices(bn)][NCH(bn)] values at the front of its array
      a.skip_cnt = (int)(T * cgr_ceil_div(H, bn_h) * (bn_h > 128 ? 2 : 1));
    }
    CgrRange prof("bwd_finalize", st);
    cgr_note_launch("bwd_finalize", st, 1);
    bwd_finalize_kernel<<<dim3((unsigned)cgr_ceil_div(H, 32), (unsigned)(a.n_col + a.n_skip)), 256, 0, st>>>(a);
    CGR_LAUNCH_CHECK();
  }
  if (side) {                                       // join: everything the side stream wrote is ordered before what follows on st
    CGR_CUDA(cudaEventRecord(side->ev[n_ev], side->s));
    CGR_CUDA(cudaStreamWaitEvent(st, side->ev[n_ev], 0));
  }
  return CGR_OK;
}


// Debug / test entry: out[M,N] = x[M,K] . w[N,K]^T + bias through the same TMA + tcgen05 FP16x3 pipeline
// (EPI_PLAIN).  workspace >= tc_linear_workspace(M, N, K).
size_t tc_linear_workspace(int64_t M, int64_t N, int64_t K) {
  const int64_t kp = round_up(K, BK);
  return cgr_align_up((size_t)M * kp * 2, 1024) * 2 + cgr_align_up((size_t)N * kp * 2, 1024) * 2 + 4096;
}

namespace {
__global__ void set_one_kernel(float* p) { *p = 1.f; }
}

int tc_linear(const float* x, int64_t M, int64_t K, int64_t ldx, const float* wgt, int64_t N, int64_t ldw,
              const float* bias, float* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CGR_CHECK_ARG(x && wgt && out && workspace && workspace_bytes >= tc_linear_workspace(M, N, K), "tc_linear: bad argument");
  const int64_t kp = round_up(K, BK);
  char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const size_t abytes = cgr_align_up((size_t)M * kp * 2, 1024), bbytes = cgr_align_up((size_t)N * kp * 2, 1024);
  __half* a_hi = (__half*)ws;
  __half* a_lo = (__half*)(ws + abytes);
  __half* b_hi = (__half*)(ws + 2 * abytes);
  __half* b_lo = (__half*)(ws + 2 * abytes + bbytes);
  float* one = (float*)(ws + 2 * abytes + 2 * bbytes);
  int* flag = (int*)(one + 16);
  CGR_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  set_one_kernel<<<1, 1, 0, st>>>(one);
  split_rows_kernel<<<(unsigned)cgr_ceil_div(M, 8), 256, 0, st>>>(x, ldx, M, (int)K, a_hi, a_lo, kp, flag, FLAG_ACT);
  split_rows_kernel<<<(unsigned)cgr_ceil_div(N, 8), 256, 0, st>>>(wgt, ldw, N, (int)K, b_hi, b_lo, kp, flag, FLAG_ACT);
  CGR_LAUNCH_CHECK();
  TcGemmParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  if ((rc = make_map(&prm.tmA_hi, a_hi, M, K, kp, TM))) return rc;
  if ((rc = make_map(&prm.tmA_lo, a_lo, M, K, kp, TM))) return rc;
  const int bn = choose_bn(cgr_ceil_div(M, TM), N);
  if ((rc = make_map(&prm.tmB_hi, b_hi, N, K, kp, bn))) return rc;
  if ((rc = make_map(&prm.tmB_lo, b_lo, N, K, kp, bn))) return rc;
  prm.num_k = (int)cgr_ceil_div(K, BK);
  prm.k_total = (int)K;
  prm.n_total = (int)N;
  prm.m_rows = (int)M;
  prm.unscale = one;
  prm.bias = bias;
  prm.out_f32 = out;
  prm.ldc = N;
  return launch_gemm<EPI_PLAIN>(prm, bn, (int)cgr_ceil_div(M, TM), true, "tc_linear", false, false, st);
}


// ------------------------------------------------------------------------------------------------
// Training path: every GEMM of the layer-wise forward / backward on tensor cores (tc_gemm2_kernel)
// ------------------------------------------------------------------------------------------------
namespace {

// max |v| of a [rows, cols] tensor: one warp per row, float4 loads when the row is 16-byte aligned
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ in, int64_t ld, int64_t rows, int cols,
                                                   unsigned int* __restrict__ amax_bits) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float m = 0.f;
  if (r < rows) {
    const float* row = in + r * ld;
    if ((ld & 3) == 0 && (cols & 3) == 0 && ((uintptr_t)in & 15) == 0) {
      for (int c = lane; c < cols / 4; c += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + c);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    } else {
      for (int c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(__ldg(row + c)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if (lane == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = wm[0];
    for (int i = 1; i < 8; ++i) b = fmaxf(b, wm[i]);
    atomicMax(amax_bits, __float_as_uint(b));       // max is order-independent: deterministic
  }
}

// fp32 [rows, cols] -> FP16 (hi, lo) with the power-of-two scale that maps amax into (2^13, 2^14];
// gradients are tiny, unscaled their lo halves would fall into the fp16 subnormal range
__global__ void __launch_bounds__(256) split_scaled_kernel(const float* __restrict__ in, int64_t ld, int64_t rows, int cols,
                                                           const unsigned int* __restrict__ amax_bits,
                                                           float* __restrict__ unscale_out, __half* __restrict__ hi,
                                                           __half* __restrict__ lo, int64_t ldo) {
  const float amax = __uint_as_float(*amax_bits);
  const float sc = (amax > 0.f && isfinite(amax)) ? exp2f(14.f - ceilf(log2f(amax))) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) *unscale_out = 1.f / sc;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* row = in + r * ld;
  if ((ld & 3) == 0 && (cols & 3) == 0 && ((uintptr_t)in & 15) == 0) {
    for (int c4 = lane; c4 < cols / 4; c4 += 32) {
      float4 v = __ldg(reinterpret_cast<const float4*>(row) + c4);
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
      uint2 ph, pl;
      ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
      pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(hi + r * ldo + 4 * c4) = ph;
      *reinterpret_cast<uint2*>(lo + r * ldo + 4 * c4) = pl;
    }
  } else {
    for (int c = lane; c < cols; c += 32) {
      __half h, l;
      split_f16(__ldg(row + c) * sc, h, l);
      hi[r * ldo + c] = h;
      lo[r * ldo + c] = l;
    }
  }
}

__global__ void splitk_reduce2_kernel(const float* __restrict__ partial, int splits, int64_t M, int64_t N,
                                      float* __restrict__ C, int64_t ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * M * N + i];   // fixed order
  C[(i / N) * ldc + (i % N)] = s;
}

struct ReduceBatchArgs {
  const float* partial[tcg2::MAX_BATCH];
  float* C[tcg2::MAX_BATCH];
  int64_t ldc[tcg2::MAX_BATCH];
};
__global__ void splitk_reduce2_batched_kernel(const ReduceBatchArgs a, int splits, int64_t M, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const float* partial = a.partial[blockIdx.y];
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * M * N + i];   // fixed order
  a.C[blockIdx.y][(i / N) * a.ldc[blockIdx.y] + (i % N)] = s;
}

int make_operand_maps(CUtensorMap* hi_map, CUtensorMap* lo_map, const TcOperand& op, int64_t mn, int64_t K) {
  // K-major: tensor [mn rows, K cols], box [128 rows, 64 cols];  MN-major: tensor [K rows, mn cols], box [64 rows, 64 cols]
  for (int h = 0; h < 2; ++h) {
    const int rc = encode_map_cached(h ? lo_map : hi_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, h ? op.lo : op.hi,
                                     (uint64_t)(op.mn_major ? mn : K), (uint64_t)(op.mn_major ? K : mn),
                                     (uint64_t)op.ld * sizeof(__half), 64u, op.mn_major ? 64u : 128u,
                                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  return CGR_OK;
}

}  // namespace

int tc_split(const float* in, int64_t ld, int64_t rows, int cols, bool scaled, __half* hi, __half* lo, int64_t ldo,
             unsigned int* amax_slot, float* unscale_slot, int* overflow, cudaStream_t st) {
  if (rows <= 0) return CGR_OK;
  CgrRange prof("tc_split", st);
  if (scaled) {
    cgr_note_launch("tc_split", st, 2);
    CGR_CUDA(cudaMemsetAsync(amax_slot, 0, sizeof(unsigned int), st));
    amax_kernel<<<(unsigned)cgr_ceil_div(rows, 8), 256, 0, st>>>(in, ld, rows, cols, amax_slot);
    split_scaled_kernel<<<(unsigned)cgr_ceil_div(rows, 8), 256, 0, st>>>(in, ld, rows, cols, amax_slot, unscale_slot, hi, lo,
                                                                         ldo);
  } else {
    cgr_note_launch("tc_split", st, 1);
    split_rows_kernel<<<(unsigned)cgr_ceil_div(rows, 8), 256, 0, st>>>(in, ld, rows, cols, hi, lo, ldo, overflow, FLAG_ACT);
  }
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int tc_splitk_choose(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = cgr_ceil_div(M, tcg2::TM) * cgr_ceil_div(N, tcg2::TN);
  const int64_t kc = cgr_ceil_div(K, tcg2::BK);
  if (tiles >= 96 || kc < 8) return 1;
  int64_t s = cgr_ceil_div(148, tiles);
  if (s > kc / 4) s = kc / 4;
  if (s > 32) s = 32;
  return s < 1 ? 1 : (int)s;
}

int tc_train_gemm(const TcOperand& A, const TcOperand& B, int64_t M, int64_t N, int64_t K, float* C, int64_t ldc,
                  const GemmEpilogue& epi, int split_k, float* partial, cudaStream_t st) {
  if (M <= 0 || N <= 0) return CGR_OK;
  tcg2::Params prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  if ((rc = make_operand_maps(&prm.tmA_hi, &prm.tmA_lo, A, M, K))) return rc;
  if ((rc = make_operand_maps(&prm.tmB_hi, &prm.tmB_lo, B, N, K))) return rc;
  const int64_t kc = cgr_ceil_div(K > 0 ? K : 1, tcg2::BK);
  if (split_k < 1) split_k = 1;
  if (split_k > 1 && !partial) { cgr_set_error("tc_train_gemm: split-K without a partial buffer"); return CGR_ERR_ARG; }
  const int64_t per = cgr_ceil_div(kc, split_k);
  split_k = (int)cgr_ceil_div(kc, per);
  prm.M = M; prm.N = N; prm.K = K;
  prm.k_chunks_per_split = per;
  prm.unscale_a = A.unscale; prm.unscale_b = B.unscale;
  prm.split_k = split_k;
  prm.C = split_k > 1 ? partial : C;
  prm.ldc = ldc;
  prm.bias = epi.bias; prm.res = epi.res; prm.ldr = epi.ldr; prm.res_scale = epi.res_scale; prm.preact = epi.preact;
  prm.act = epi.act; prm.dropout_p = epi.dropout_p; prm.seed = epi.seed; prm.layer = epi.layer;
  static bool attr_done = false;
  if (!attr_done) {
    CGR_CUDA(cudaFuncSetAttribute(tcg2::tc_gemm2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg2::SMEM_BYTES));
    CGR_CUDA(cudaFuncSetAttribute(tcg2::tc_gemm2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg2::SMEM_BYTES));
    CGR_CUDA(cudaFuncSetAttribute(tcg2::tc_gemm2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg2::SMEM_BYTES));
    CGR_CUDA(cudaFuncSetAttribute(tcg2::tc_gemm2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg2::SMEM_BYTES));
    attr_done = true;
  }
  const char* name = epi.tag ? epi.tag : "tc_gemm2";
  CgrRange prof(name, st);
  cgr_note_launch(name, st, split_k > 1 ? 2 : 1);
  dim3 grid((unsigned)cgr_ceil_div(N, tcg2::TN), (unsigned)cgr_ceil_div(M, tcg2::TM), (unsigned)split_k);
  if (!A.mn_major && !B.mn_major) tcg2::tc_gemm2_kernel<false, false><<<grid, tcg2::THREADS, tcg2::SMEM_BYTES, st>>>(prm);
  else if (!A.mn_major && B.mn_major) tcg2::tc_gemm2_kernel<false, true><<<grid, tcg2::THREADS, tcg2::SMEM_BYTES, st>>>(prm);
  else if (A.mn_major && B.mn_major) tcg2::tc_gemm2_kernel<true, true><<<grid, tcg2::THREADS, tcg2::SMEM_BYTES, st>>>(prm);
  else tcg2::tc_gemm2_kernel<true, false><<<grid, tcg2::THREADS, tcg2::SMEM_BYTES, st>>>(prm);
  if (split_k > 1)
    splitk_reduce2_kernel<<<(unsigned)cgr_ceil_div(M * N, 256), 256, 0, st>>>(partial, split_k, M, N, C, ldc);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

// Split-K for a batch of n same-shape reductions: one wave of CTAs (148 SMs, one CTA each) when K is short -- every CTA
// pays a fixed prologue + epilogue -- and up to four waves of >= 32-chunk CTAs when K is long.
int tc_splitk_batched(int n, int64_t M, int64_t N, int64_t K) {
  const int64_t base = (int64_t)n * cgr_ceil_div(M, tcg2::TM) * cgr_ceil_div(N, tcg2::TN);
  const int64_t kc = cgr_ceil_div(K > 0 ? K : 1, tcg2::BK);
  int64_t s = 148 / base;
  if (s < 1) s = 1;
  if (kc / s >= 64) {
    int64_t s2 = 592 / base;
    if (s2 > kc / 32) s2 = kc / 32;
    if (s2 > s) s = s2;
  }
  if (s > kc) s = kc;
  if (s > 64) s = 64;
  return (int)(s < 1 ? 1 : s);
}

// Up to 8 reductions C_i[M,N] = A_i^T B_i of ONE shape (both operands [K rows, MN cols]) in one GEMM launch plus one
// split-K reduction launch.  `partial` holds n * split_k * M * N floats.
int tc_train_gemm_batched_mn(const TcOperand* A, const TcOperand* B, float* const* C, const int64_t* ldc, int n,
                             int64_t M, int64_t N, int64_t K, int split_k, float* partial, const char* name,
                             cudaStream_t st) {
  CGR_CHECK_ARG(n >= 1 && n <= tcg2::MAX_BATCH && M > 0 && N > 0 && K > 0, "tc_train_gemm_batched_mn: bad argument");
  static tcg2::BatchParams bp;       // large (several KB): built in place; launches copy it
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  memset(&bp, 0, sizeof(bp));
  const int64_t kc = cgr_ceil_div(K, tcg2::BK);
  if (split_k < 1) split_k = 1;
  const int64_t per = cgr_ceil_div(kc, split_k);
  split_k = (int)cgr_ceil_div(kc, per);
  CGR_CHECK_ARG(split_k == 1 || partial, "tc_train_gemm_batched_mn: partial buffer missing");
  ReduceBatchArgs ra;
  memset(&ra, 0, sizeof(ra));
  int rc;
  for (int i = 0; i < n; ++i) {
    tcg2::Params& prm = bp.prob[i];
    CGR_CHECK_ARG(A[i].mn_major && B[i].mn_major, "tc_train_gemm_batched_mn: operands must be [K, MN]");
    if ((rc = make_operand_maps(&prm.tmA_hi, &prm.tmA_lo, A[i], M, K))) return rc;
    if ((rc = make_operand_maps(&prm.tmB_hi, &prm.tmB_lo, B[i], N, K))) return rc;
    prm.M = M; prm.N = N; prm.K = K;
    prm.k_chunks_per_split = per;
    prm.unscale_a = A[i].unscale; prm.unscale_b = B[i].unscale;
    prm.split_k = split_k;
    prm.act = CGR_ACT_IDENTITY;
    if (split_k > 1) {
      prm.C = partial + (size_t)i * split_k * M * N;
      prm.ldc = N;
    } else {                                     // no split: the GEMM writes the gradient directly
      prm.C = C[i];
      prm.ldc = ldc[i];
    }
    ra.partial[i] = prm.C; ra.C[i] = C[i]; ra.ldc[i] = ldc[i];
  }
  bp.n_prob = n;
  static bool attr_done = false;
  if (!attr_done) {
    CGR_CUDA(cudaFuncSetAttribute(tcg2::tc_gemm2_batched_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg2::SMEM_BYTES));
    attr_done = true;
  }
  CgrRange prof(name, st);
  cgr_note_launch(name, st, split_k > 1 ? 2 : 1);
  dim3 grid((unsigned)cgr_ceil_div(N, tcg2::TN), (unsigned)cgr_ceil_div(M, tcg2::TM), (unsigned)(split_k * n));
  tcg2::tc_gemm2_batched_kernel<true, true><<<grid, tcg2::THREADS, tcg2::SMEM_BYTES, st>>>(bp);
  if (split_k > 1)
    splitk_reduce2_batched_kernel<<<dim3((unsigned)cgr_ceil_div(M * N, 256), (unsigned)n), 256, 0, st>>>(ra, split_k, M, N);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

// prepared-weight accessor: matrix `mat` (0 = [W_x; W_ox], 1..depth = convs, depth+1 = W_os), rows from `row0`
TcOperand tc_weight_operand(const cgr_params_t* p, const void* wbuf, int mat, int64_t row0, bool mn_major) {
  const WLayout wl = wlayout(p);
  const char* b = (const char*)wbuf;
  TcOperand op;
  op.hi = (const __half*)(b + wl.off_hi[mat]) + row0 * wl.ld[mat];
  op.lo = (const __half*)(b + wl.off_lo[mat]) + row0 * wl.ld[mat];
  op.ld = wl.ld[mat];
  op.unscale = (const float*)(b + wl.off_unscale) + mat;
  op.mn_major = mn_major;
  return op;
}

// Test entry: C[M,N] = op(A) op(B)^T style contraction on fp32 inputs.  a_mn: A given as [K, M]; b_mn: B given as [K, N].
size_t tc_gemm2_test_workspace(int64_t M, int64_t N, int64_t K) {
  const int64_t a_el = (M > K ? M : K) * round_up(M > K ? M : K, 64), b_el = (N > K ? N : K) * round_up(N > K ? N : K, 64);
  return (size_t)(a_el + b_el) * 2 * sizeof(__half) + (size_t)32 * M * N * sizeof(float) + 8192;
}
int tc_gemm2_test(const float* A, const float* B, int64_t M, int64_t N, int64_t K, int a_mn, int b_mn, float* C,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CGR_CHECK_ARG(A && B && C && workspace && workspace_bytes >= tc_gemm2_test_workspace(M, N, K), "tc_gemm2_test: bad argument");
  char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const int64_t a_rows = a_mn ? K : M, a_cols = a_mn ? M : K, b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  const int64_t a_ld = round_up(a_cols, 64), b_ld = round_up(b_cols, 64);
  const int64_t a_el = (M > K ? M : K) * round_up(M > K ? M : K, 64), b_el = (N > K ? N : K) * round_up(N > K ? N : K, 64);
  __half* a_hi = (__half*)ws;
  __half* a_lo = a_hi + a_el;
  __half* b_hi = a_lo + a_el;
  __half* b_lo = b_hi + b_el;
  float* partial = (float*)(b_lo + b_el);
  unsigned int* slots = (unsigned int*)(partial + 32 * M * N);
  float* us = (float*)(slots + 8);
  int rc;
  if ((rc = tc_split(A, a_cols, a_rows, (int)a_cols, true, a_hi, a_lo, a_ld, slots, us, nullptr, st))) return rc;
  if ((rc = tc_split(B, b_cols, b_rows, (int)b_cols, true, b_hi, b_lo, b_ld, slots + 1, us + 1, nullptr, st))) return rc;
  TcOperand oa{a_hi, a_lo, a_ld, us, a_mn != 0}, ob{b_hi, b_lo, b_ld, us + 1, b_mn != 0};
  GemmEpilogue e;
  return tc_train_gemm(oa, ob, M, N, K, C, N, e, tc_splitk_choose(M, N, K), partial, st);
}
