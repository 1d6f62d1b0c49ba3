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
#include <string.h>

#include "umma.cuh"

namespace {

constexpr int TM = 128;                 // rows of a tile (UMMA M)
constexpr int BK = 64;                  // fp16 elements per k-chunk = one 128-byte swizzle row
constexpr int BN = 80;                  // output columns per CTA (UMMA N), multiple of 16
constexpr int BNP = BN + 4;             // padded row of the fp32 staging tiles (conflict-free float4 rows)
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * BK * 2;    // 16 KB
constexpr int B_BYTES = BN * BK * 2;    // 10 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int TMEM_COLS = 128;          // power of two >= BN
constexpr int THREADS = 256;
constexpr int AUX_BYTES = 2048;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;   // + alignment slack
static_assert(2 * TM * BNP * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit in the pipeline buffers");

enum { EPI_PLAIN = 0, EPI_BOND = 1, EPI_READOUT = 2 };

struct TcGemmParams {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  int num_k;                    // k-chunks of BK
  int n_total;                  // real output columns
  int m_rows;                   // real rows (EPI_PLAIN)
  const float* unscale;         // device scalar 1 / (scale_A * scale_W)
  const float* bias;            // [n_total] or null
  // EPI_PLAIN
  float* out_f32;
  int64_t ldc;
  // tile-local epilogues
  const int32_t* tile_info;     // [T][8]: ebase, ecount, abase, acount, rx0, rxcount, 0, 0
  const int32_t* in_ptr;
  const int32_t* in_idx;
  const int32_t* src;
  const int32_t* atom_ptr;
  const float* skip;            // device scalar or null (=1)
  const float* h0;              // [T*128, H] fp32 (tile-packed rows)
  int act;
  float dropout_p;
  uint64_t seed;
  uint32_t layer;
  __half* o_hi;                 // next operand, [T*128, ldo]
  __half* o_lo;
  int64_t ldo;
  const float* Q;               // readout: x W_ox^T + b_o, [N, ldq]
  int64_t ldq;
  const float* w_ffn;
  float* partial_out;           // [n_slices, B]
  int64_t n_rxn;
  int* overflow;                // sticky flag: an activation left the fp16 range
};

struct Aux {                    // small per-CTA shared state, lives after the pipeline buffers
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full;
  uint32_t tmem_base;
  int32_t info[8];
  uint16_t ptr_l[TM + 2];       // local CSR offsets of the tile's atoms
  uint8_t src_l[TM];            // local source atom of each bond row
  uint8_t idx_l[TM];            // local bond ids grouped by target atom
  float tat[TM];                // readout: per-atom dot with w_ffn
};
static_assert(sizeof(Aux) <= AUX_BYTES, "Aux too large");

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = umma::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  Aux* aux = reinterpret_cast<Aux*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, slice = blockIdx.y;
  const int n0 = slice * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->empty[s]), 1);
    }
    umma::mbar_init(umma::smem_u32(&aux->tmem_full), 1);
    umma::mbar_fence_init();
    umma::tma_prefetch_desc(&p.tmA_hi);
    umma::tma_prefetch_desc(&p.tmA_lo);
    umma::tma_prefetch_desc(&p.tmB_hi);
    umma::tma_prefetch_desc(&p.tmB_lo);
  }
  if (warp == 1) {
    umma::tmem_alloc(umma::smem_u32(&aux->tmem_base), TMEM_COLS);
    umma::tmem_relinquish();
  }
  if (EPI != EPI_PLAIN && threadIdx.x < 8) aux->info[threadIdx.x] = __ldg(p.tile_info + (int64_t)tile * 8 + threadIdx.x);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = aux->tmem_base;

  // ------------------------------------------------------------------ main loop (warp-specialised)
  if (warp == 0) {
    // TMA producer: one elected lane streams A (hi, lo) and B (hi, lo) k-chunks through the ring
    for (int kc = 0; kc < p.num_k; ++kc) {
      const int s = kc % STAGES;
      const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
      if (lane == 0) {
        umma::mbar_wait(umma::smem_u32(&aux->empty[s]), ph ^ 1u);
        const uint32_t full = umma::smem_u32(&aux->full[s]);
        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
        umma::mbar_arrive_expect_tx(full, STAGE_BYTES);
        umma::tma_load_2d(&p.tmA_hi, full, st, kc * BK, tile * TM);
        umma::tma_load_2d(&p.tmA_lo, full, st + A_BYTES, kc * BK, tile * TM);
        umma::tma_load_2d(&p.tmB_hi, full, st + 2 * A_BYTES, kc * BK, n0);
        umma::tma_load_2d(&p.tmB_lo, full, st + 2 * A_BYTES + B_BYTES, kc * BK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // MMA issuer: one lane issues 3 tcgen05.mma per 16-wide k-step (hi.hi + lo.hi + hi.lo)
    constexpr uint32_t idesc = umma::idesc_f16_f32(TM, BN);
    for (int kc = 0; kc < p.num_k; ++kc) {
      const int s = kc % STAGES;
      const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
      if (lane == 0) {
        umma::mbar_wait(umma::smem_u32(&aux->full[s]), ph);
        umma::tc_fence_after_sync();
        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
        const uint64_t da_hi = umma::smem_desc_k_sw128(st);
        const uint64_t da_lo = umma::smem_desc_k_sw128(st + A_BYTES);
        const uint64_t db_hi = umma::smem_desc_k_sw128(st + 2 * A_BYTES);
        const uint64_t db_lo = umma::smem_desc_k_sw128(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 32 >> 4);          // 16 fp16 = 32 bytes along K inside the swizzle row
          umma::mma_f16_ss(tmem, da_lo + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
          umma::mma_f16_ss(tmem, da_hi + adv, db_lo + adv, idesc, 1u);
          umma::mma_f16_ss(tmem, da_hi + adv, db_hi + adv, idesc, 1u);
        }
        umma::mma_commit(umma::smem_u32(&aux->empty[s]));         // frees the stage when these MMAs retire
        if (kc == p.num_k - 1) umma::mma_commit(umma::smem_u32(&aux->tmem_full));
      }
      __syncwarp();
    }
  } else if (EPI != EPI_PLAIN) {
    // the other warps stage the tile's index rows into shared memory while the GEMM runs
    const int ebase = aux->info[0], ecount = aux->info[1], abase = aux->info[2], acount = aux->info[3];
    for (int j = threadIdx.x - 64; j < ecount; j += THREADS - 64) {
      aux->src_l[j] = (uint8_t)(__ldg(p.src + ebase + j) - abase);
      aux->idx_l[j] = (uint8_t)(__ldg(p.in_idx + ebase + j) - ebase);
    }
    for (int v = threadIdx.x - 64; v <= acount; v += THREADS - 64)
      aux->ptr_l[v] = (uint16_t)(__ldg(p.in_ptr + abase + v) - ebase);
  }

  // ------------------------------------------------------------------ epilogue (all 8 warps)
  umma::mbar_wait(umma::smem_u32(&aux->tmem_full), 0);
  umma::tc_fence_after_sync();
  float* y_s = reinterpret_cast<float*>(smem);                   // [TM][BNP], aliases the drained pipeline
  float* a_s = y_s + TM * BNP;                                   // [TM][BNP]
  {
    const float us = __ldg(p.unscale);
    const int q = warp & 3, half = warp >> 2;                    // TMEM lane quarter / column half
    const int row = q * 32 + lane;
    constexpr int COLS_PER_WARP = BN / 2;                        // 40
#pragma unroll
    for (int cc = 0; cc < COLS_PER_WARP; cc += 8) {
      const int c = half * COLS_PER_WARP + cc;
      float v[8];
      umma::tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      umma::tmem_ld_wait();
      float4* dst = reinterpret_cast<float4*>(y_s + row * BNP + c);
      dst[0] = make_float4(v[0] * us, v[1] * us, v[2] * us, v[3] * us);
      dst[1] = make_float4(v[4] * us, v[5] * us, v[6] * us, v[7] * us);
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();

  if (EPI == EPI_PLAIN) {
    // rows are dense (atoms): out = y + bias, written as coalesced rows
    for (int r = warp; r < TM; r += THREADS / 32) {
      const int64_t row = (int64_t)tile * TM + r;
      if (row >= p.m_rows) break;
      for (int c = lane; c < BN; c += 32) {
        const int n = n0 + c;
        if (n < p.n_total) p.out_f32[row * p.ldc + n] = y_s[r * BNP + c] + (p.bias ? __ldg(p.bias + n) : 0.f);
      }
    }
  } else if (EPI == EPI_BOND) {
    const int ebase = aux->info[0], ecount = aux->info[1], acount = aux->info[3];
    // a[v] = sum_{k in in(v)} y[k]   (ascending bond id, the reference's accumulation order)
    for (int v = warp; v < acount; v += THREADS / 32) {
      const int pb = aux->ptr_l[v], pe = aux->ptr_l[v + 1];
      for (int c = lane; c < BN; c += 32) {
        float a = 0.f;
        for (int q = pb; q < pe; ++q) a += y_s[(int)aux->idx_l[q] * BNP + c];
        a_s[v * BNP + c] = a;
      }
    }
    __syncthreads();
    const float skip = p.skip ? __ldg(p.skip) : 1.f;
    const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    const int H = p.n_total;
    bool ovf = false;
    // z[e] = a[src e] - y[e^1] + b + skip*h0[e];  h' = dropout(act(z));  written as the FP16 (hi, lo) operand
    for (int j = warp; j < ecount; j += THREADS / 32) {
      const int64_t r = (int64_t)tile * TM + j;
      const float* arow = a_s + (int)aux->src_l[j] * BNP;
      const float* yrev = y_s + (j ^ 1) * BNP;
      for (int c2 = lane; c2 < BN / 2; c2 += 32) {
        const int c = 2 * c2, n = n0 + c;
        if (n >= H) continue;
        const float2 h0v = *reinterpret_cast<const float2*>(p.h0 + r * H + n);
        float z0 = arow[c] - yrev[c] + __ldg(p.bias + n) + skip * h0v.x;
        float z1 = arow[c + 1] - yrev[c + 1] + __ldg(p.bias + n + 1) + skip * h0v.y;
        z0 = cgr_act(z0, p.act);
        z1 = cgr_act(z1, p.act);
        if (p.dropout_p > 0.f) {
          const uint64_t idx = (uint64_t)(ebase + j) * (uint64_t)H + (uint64_t)n;
          z0 = cgr_dropout_keep(p.seed, p.layer, idx, p.dropout_p) ? z0 * keep_scale : 0.f;
          z1 = cgr_dropout_keep(p.seed, p.layer, idx + 1, p.dropout_p) ? z1 * keep_scale : 0.f;
        }
        ovf |= (fabsf(z0) > 60000.f) | (fabsf(z1) > 60000.f);
        __half h0h, h0l, h1h, h1l;
        split_f16(z0, h0h, h0l);
        split_f16(z1, h1h, h1l);
        *reinterpret_cast<__half2*>(p.o_hi + r * p.ldo + n) = __halves2half2(h0h, h1h);
        *reinterpret_cast<__half2*>(p.o_lo + r * p.ldo + n) = __halves2half2(h0l, h1l);
      }
    }
    if (ovf) atomicOr(p.overflow, 1);
  } else {
    // readout: hv[v] = act(Q[v] + sum_{k in in(v)} y[k]);  t[v] = hv[v] . w_f (this CTA's columns)
    const int abase = aux->info[2], acount = aux->info[3], rx0 = aux->info[4], rxcount = aux->info[5];
    const int H = p.n_total;
    for (int v = warp; v < acount; v += THREADS / 32) {
      const int pb = aux->ptr_l[v], pe = aux->ptr_l[v + 1];
      float t = 0.f;
      for (int c = lane; c < BN; c += 32) {
        const int n = n0 + c;
        if (n >= H) continue;
        float a = 0.f;
        for (int q = pb; q < pe; ++q) a += y_s[(int)aux->idx_l[q] * BNP + c];
        const float zv = __ldg(p.Q + (int64_t)(abase + v) * p.ldq + n) + a;
        t = fmaf(cgr_act(zv, p.act), __ldg(p.w_ffn + n), t);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) aux->tat[v] = t;
    }
    __syncthreads();
    for (int rx = threadIdx.x; rx < rxcount; rx += THREADS) {
      const int b = rx0 + rx;
      const int v0 = __ldg(p.atom_ptr + b) - abase, v1 = __ldg(p.atom_ptr + b + 1) - abase;
      float s = 0.f;
      for (int v = v0; v < v1; ++v) s += aux->tat[v];            // ascending atom id
      p.partial_out[(int64_t)slice * p.n_rxn + b] = s;
    }
  }

  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// small SIMT kernels around the GEMMs
// ------------------------------------------------------------------------------------------------

// fp32 rows -> (hi, lo) fp16 rows; one warp per row
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ in, int64_t ld, int64_t rows, int cols,
                                                         __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldo,
                                                         int* __restrict__ overflow) {
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
  if (ovf) atomicOr(overflow, 1);
}

// h0 = act(P'[src e] + ea[e] . W_e^T) on tile-packed rows; writes the fp32 copy (skip input) and the split operand
__global__ void __launch_bounds__(128) tc_edge_init_kernel(const float* __restrict__ PQ, int64_t ldpq,
                                                           const float* __restrict__ ea, const int32_t* __restrict__ src,
                                                           const float* __restrict__ w_init,
                                                           const int32_t* __restrict__ tile_info, int fa, int fb, int H,
                                                           int act, float* __restrict__ h0, __half* __restrict__ o_hi,
                                                           __half* __restrict__ o_lo, int64_t ldo,
                                                           int* __restrict__ overflow) {
  extern __shared__ float ea_s[];                  // [4][fb]
  const int tile = blockIdx.y;
  const int ebase = __ldg(tile_info + tile * 8), ecount = __ldg(tile_info + tile * 8 + 1);
  const int j0 = blockIdx.x * 4;
  if (j0 >= ecount) return;
  for (int i = threadIdx.x; i < 4 * fb; i += blockDim.x) {
    const int j = j0 + i / fb;
    ea_s[i] = j < ecount ? __ldg(ea + (int64_t)(ebase + j) * fb + (i % fb)) : 0.f;
  }
  __syncthreads();
  const int ld = fa + fb;
  bool ovf = false;
  for (int n = threadIdx.x; n < H; n += blockDim.x) {
    const float* w = w_init + (int64_t)n * ld + fa;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < fb; ++k) {
      const float wk = __ldg(w + k);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(ea_s[i * fb + k], wk, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i;
      if (j >= ecount) break;
      const int64_t r = (int64_t)tile * TM + j;
      const float v = cgr_act(__ldg(PQ + (int64_t)__ldg(src + ebase + j) * ldpq + n) + acc[i], act);
      ovf |= fabsf(v) > 60000.f;
      h0[r * H + n] = v;
      __half h, l;
      split_f16(v, h, l);
      o_hi[r * ldo + n] = h;
      o_lo[r * ldo + n] = l;
    }
  }
  if (ovf) atomicOr(overflow, 1);
}

__global__ void tc_finalize_kernel(const float* __restrict__ partial, int n_slices, int64_t B,
                                   const float* __restrict__ b_ffn, float* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int i = 0; i < n_slices; ++i) s += partial[(int64_t)i * B + b];     // fixed order
  out[b] = s + __ldg(b_ffn);
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
  int64_t ldo[MAX_SEG];
};

__global__ void __launch_bounds__(256) prep_amax_kernel(const PrepArgs a) {
  const Seg s = a.seg[blockIdx.y];
  const int64_t total = (int64_t)s.rows * s.cols;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(s.src + (i / s.cols) * s.ld + (i % s.cols))));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(a.amax_bits + s.mat, __float_as_uint(m));   // max is order-independent
}

__device__ __forceinline__ float weight_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  return exp2f(13.f - ceilf(log2f(amax)));        // amax * scale in (2^12, 2^13]
}

__global__ void __launch_bounds__(256) prep_split_kernel(const PrepArgs a) {
  const Seg s = a.seg[blockIdx.y];
  const float sc = weight_scale(__uint_as_float(a.amax_bits[s.mat]));
  if (blockIdx.x == 0 && threadIdx.x == 0 && s.row0 == 0) a.unscale[s.mat] = 1.f / sc;
  const int64_t total = (int64_t)s.rows * s.cols;
  __half* hi = a.hi[s.mat];
  __half* lo = a.lo[s.mat];
  const int64_t ldo = a.ldo[s.mat];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / s.cols, c = i % s.cols;
    __half h, l;
    split_f16(__ldg(s.src + r * s.ld + c) * sc, h, l);
    hi[(s.row0 + r) * ldo + c] = h;
    lo[(s.row0 + r) * ldo + c] = l;
  }
}

__global__ void concat_bias_kernel(const float* __restrict__ b0, const float* __restrict__ b1, int H,
                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H) {
    out[i] = __ldg(b0 + i);
    out[H + i] = __ldg(b1 + i);
  }
}

// greedy packing of consecutive whole reactions into 128-row tiles (bonds and atoms both <= 128)
__global__ void tile_plan_kernel(const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ atom_ptr, int64_t B,
                                 int32_t* __restrict__ tile_info, int32_t* __restrict__ status) {
  extern __shared__ int32_t sh[];                  // chunks of (atom offset, bond offset)
  constexpr int CH = 2048;
  int32_t* a_s = sh;
  int32_t* e_s = sh + CH + 1;
  int t = -1, used_e = TM + 1, used_a = TM + 1, ok = 1;
  int cur[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t g0 = 0; g0 < B; g0 += CH) {
    const int n = (int)((B - g0) < CH ? (B - g0) : CH);
    __syncthreads();
    for (int i = threadIdx.x; i <= n; i += blockDim.x) {
      const int32_t a = __ldg(atom_ptr + g0 + i);
      a_s[i] = a;
      e_s[i] = __ldg(in_ptr + a);                  // bonds of a reaction are the in-bonds of its atoms
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 0; i < n; ++i) {
        const int ne = e_s[i + 1] - e_s[i], na = a_s[i + 1] - a_s[i];
        if (ne > TM || na > TM || ne <= 0 || na <= 0 || (ne & 1)) { ok = 0; continue; }
        if (used_e + ne > TM || used_a + na > TM) {
          if (t >= 0)
            for (int k = 0; k < 6; ++k) tile_info[t * 8 + k] = cur[k];
          ++t;
          cur[0] = e_s[i]; cur[1] = 0; cur[2] = a_s[i]; cur[3] = 0; cur[4] = (int)(g0 + i); cur[5] = 0;
          used_e = 0; used_a = 0;
        }
        used_e += ne; used_a += na;
        cur[1] = used_e; cur[3] = used_a; cur[5] += 1;
      }
    }
  }
  if (threadIdx.x == 0) {
    if (t >= 0)
      for (int k = 0; k < 6; ++k) tile_info[t * 8 + k] = cur[k];
    status[0] = t + 1;
    status[1] = ok;
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

// fp16 matrix [rows, cols] with row stride `ld` elements; box = [box_rows, 64 cols]; 128-byte swizzle;
// out-of-bounds elements (K tail, row tail) are zero-filled by the TMA unit.
int make_map(CUtensorMap* tm, const __half* basep, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { cgr_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CGR_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)basep, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { cgr_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return CGR_ERR_ARG; }
  return CGR_OK;
}

int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

struct WLayout {                 // prepared-weight buffer: [amax | unscale | bias_cat | matrices (hi, lo)...]
  int n_mat;
  int64_t kp_x, kp_h;
  size_t off_amax, off_unscale, off_bias, off_hi[MAX_SEG], off_lo[MAX_SEG], total;
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
  for (int m = 0; m < w.n_mat; ++m) {
    w.rows[m] = m == 0 ? 2 * H : H;
    w.ld[m] = m == 0 ? w.kp_x : w.kp_h;
    const size_t bytes = cgr_align_up((size_t)w.rows[m] * w.ld[m] * sizeof(__half), 1024);
    w.off_hi[m] = off; off += bytes;
    w.off_lo[m] = off; off += bytes;
  }
  w.total = off;
  return w;
}

template <int EPI>
int launch_gemm(const TcGemmParams& prm, int m_tiles, int n_slices, const char* name, cudaStream_t st) {
  static bool attr_done = false;      // benign race: the attribute is idempotent
  if (!attr_done) {
    CGR_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  CgrRange prof(name, st);
  cgr_note_launch(name, st, 1);
  tc_gemm_kernel<EPI><<<dim3((unsigned)m_tiles, (unsigned)n_slices), THREADS, SMEM_BYTES, st>>>(prm);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------

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
    a.ldo[m] = w.ld[m];
  }
  int ns = 0;
  a.seg[ns++] = Seg{p->w_init, fa + fb, H, fa, 0, 0};              // W_x   = edge_init.weight[:, :Fa]
  a.seg[ns++] = Seg{p->w_e2n, fa + H, H, fa, 0, H};                // W_ox  = edge_to_node.weight[:, :Fa]
  for (int l = 0; l < p->depth; ++l) a.seg[ns++] = Seg{p->w_conv[l], H, H, H, 1 + l, 0};
  a.seg[ns++] = Seg{p->w_e2n + fa, fa + H, H, H, p->depth + 1, 0};  // W_os  = edge_to_node.weight[:, Fa:]
  a.n_seg = ns;
  CgrRange prof("tc_prep_weights", st);
  cgr_note_launch("tc_prep_weights", st, 3);
  CGR_CUDA(cudaMemsetAsync(a.amax_bits, 0, MAX_SEG * sizeof(unsigned int), st));
  prep_amax_kernel<<<dim3(64, (unsigned)ns), 256, 0, st>>>(a);
  prep_split_kernel<<<dim3(64, (unsigned)ns), 256, 0, st>>>(a);
  concat_bias_kernel<<<(unsigned)cgr_ceil_div(H, 256), 256, 0, st>>>(p->b_init, p->b_e2n, H, (float*)(b + w.off_bias));
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int tc_plan_build(const int32_t* in_ptr, const int32_t* atom_ptr, const int32_t* src, const int32_t* dst,
                  int64_t n_rxn, int32_t* tile_info, int32_t* status, cudaStream_t st) {
  CGR_CHECK_ARG(in_ptr && atom_ptr && tile_info && status && n_rxn > 0, "tc_plan_build: bad argument");
  (void)src; (void)dst;            // endpoints are validated by tc_plan_check once the tile count is known
  cgr_note_launch("tc_plan", st, 1);
  tile_plan_kernel<<<1, 256, (2 * 2048 + 2) * sizeof(int32_t), st>>>(in_ptr, atom_ptr, n_rxn, tile_info, status);
  CGR_LAUNCH_CHECK();
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
  TcWs w;
  const int64_t H = p->hidden, N = g->n_atoms, T = g->n_tiles, B = g->n_rxn;
  w.kp_x = round_up(p->fa, BK);
  w.kp_h = round_up(H, BK);
  w.rows_pad = T * TM;
  w.n_slices = (int)cgr_ceil_div(H, BN);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += cgr_align_up(bytes, 1024); return o; };
  w.off_w = take(need_w ? tc_weights_bytes(p) : 0);
  w.off_xhi = take((size_t)N * w.kp_x * sizeof(__half));
  w.off_xlo = take((size_t)N * w.kp_x * sizeof(__half));
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

size_t tc_forward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int training) {
  (void)training;
  if (!g->tile_info || g->n_tiles <= 0) return 0;
  return tc_ws(p, g, p->tc_weights == nullptr).total;
}

int tc_gnn_forward(const cgr_params_t* p, const cgr_graph_t* g, float* out, cgr_saved_t* saved, int training,
                   uint64_t seed, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CGR_CHECK_ARG(!saved, "tcgen05 engine: the training (activation-saving) forward is not available yet; use engine simt");
  CGR_CHECK_ARG(g->tile_info && g->n_tiles > 0, "tcgen05 engine needs a tile plan (reactions of <= 128 bonds)");
  CGR_CHECK_ARG(p->hidden % 2 == 0, "tcgen05 engine needs an even hidden size");
  CGR_CHECK_ARG(p->depth + 3 <= MAX_SEG, "tcgen05 engine supports depth <= %d", MAX_SEG - 3);
  const bool need_w = p->tc_weights == nullptr;
  const TcWs w = tc_ws(p, g, need_w);
  CGR_CHECK_ARG(workspace && workspace_bytes >= w.total, "tc_gnn_forward: workspace too small");
  char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const int H = p->hidden, fa = p->fa, fb = p->fb, d = p->depth;
  const int64_t N = g->n_atoms, T = g->n_tiles, B = g->n_rxn;
  int rc;

  char* wbuf = need_w ? ws + w.off_w : (char*)p->tc_weights;
  if (need_w) {
    rc = tc_prepare_weights(p, wbuf, tc_weights_bytes(p), st);
    if (rc) return rc;
  }
  const WLayout wl = wlayout(p);
  const float* unscale = (const float*)(wbuf + wl.off_unscale);
  const float* bias_cat = (const float*)(wbuf + wl.off_bias);
  auto w_hi = [&](int m) { return (const __half*)(wbuf + wl.off_hi[m]); };
  auto w_lo = [&](int m) { return (const __half*)(wbuf + wl.off_lo[m]); };

  __half* x_hi = (__half*)(ws + w.off_xhi);
  __half* x_lo = (__half*)(ws + w.off_xlo);
  float* PQ = (float*)(ws + w.off_pq);
  float* h0 = (float*)(ws + w.off_h0);
  __half* h_hi[2] = {(__half*)(ws + w.off_hhi[0]), (__half*)(ws + w.off_hhi[1])};
  __half* h_lo[2] = {(__half*)(ws + w.off_hlo[0]), (__half*)(ws + w.off_hlo[1])};
  float* partial = (float*)(ws + w.off_partial);
  int* flag = (int*)(ws + w.off_flag);
  CGR_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));

  // 1. x -> (hi, lo)
  {
    CgrRange prof("tc_split_x", st);
    cgr_note_launch("tc_split_x", st, 1);
    split_rows_kernel<<<(unsigned)cgr_ceil_div(N, 8), 256, 0, st>>>(g->x, fa, N, fa, x_hi, x_lo, w.kp_x, flag);
    CGR_LAUNCH_CHECK();
  }
  // 2. per-atom projections [P' | Q'] = x [W_x ; W_ox]^T + [b_i | b_o]   (GNN.py:86 and :106-107, x part)
  {
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    if ((rc = make_map(&prm.tmA_hi, x_hi, N, fa, w.kp_x, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, x_lo, N, fa, w.kp_x, TM))) return rc;
    if ((rc = make_map(&prm.tmB_hi, w_hi(0), 2 * H, fa, wl.ld[0], BN))) return rc;
    if ((rc = make_map(&prm.tmB_lo, w_lo(0), 2 * H, fa, wl.ld[0], BN))) return rc;
    prm.num_k = (int)cgr_ceil_div(fa, BK);
    prm.n_total = 2 * H;
    prm.m_rows = (int)N;
    prm.unscale = unscale + 0;
    prm.bias = bias_cat;
    prm.out_f32 = PQ;
    prm.ldc = 2 * H;
    rc = launch_gemm<EPI_PLAIN>(prm, (int)cgr_ceil_div(N, TM), (int)cgr_ceil_div(2 * H, BN), "tc_atom_proj", st);
    if (rc) return rc;
  }
  // 3. edge initialisation on tile-packed rows
  {
    CgrRange prof("tc_edge_init", st);
    cgr_note_launch("tc_edge_init", st, 1);
    const size_t smem = (size_t)(4 * (fb > 0 ? fb : 1)) * sizeof(float);
    tc_edge_init_kernel<<<dim3(TM / 4, (unsigned)T), 128, smem, st>>>(PQ, 2 * H, g->edge_attr, g->src, p->w_init,
                                                                      g->tile_info, fa, fb, H, p->act, h0, h_hi[0],
                                                                      h_lo[0], w.kp_h, flag);
    CGR_LAUNCH_CHECK();
  }
  // 4. message passing layers: one fused kernel each
  for (int l = 0; l < d; ++l) {
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    const int in = l & 1, ob = in ^ 1;
    if ((rc = make_map(&prm.tmA_hi, h_hi[in], w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, h_lo[in], w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmB_hi, w_hi(1 + l), H, H, wl.ld[1 + l], BN))) return rc;
    if ((rc = make_map(&prm.tmB_lo, w_lo(1 + l), H, H, wl.ld[1 + l], BN))) return rc;
    prm.num_k = (int)cgr_ceil_div(H, BK);
    prm.n_total = H;
    prm.unscale = unscale + 1 + l;
    prm.bias = p->b_conv[l];
    prm.tile_info = g->tile_info;
    prm.in_ptr = g->in_ptr; prm.in_idx = g->in_idx; prm.src = g->src; prm.atom_ptr = g->atom_ptr;
    prm.skip = p->use_skip ? p->skip[l] : nullptr;
    prm.h0 = h0;
    prm.act = p->act;
    prm.dropout_p = (training && p->host_dropout_p) ? p->host_dropout_p[l] : 0.f;
    prm.seed = seed; prm.layer = (uint32_t)l;
    prm.o_hi = h_hi[ob]; prm.o_lo = h_lo[ob]; prm.ldo = w.kp_h;
    prm.overflow = flag;
    rc = launch_gemm<EPI_BOND>(prm, (int)T, w.n_slices, "bond_layer", st);
    if (rc) return rc;
  }
  // 5. readout + pooling + FFN
  {
    TcGemmParams prm;
    memset(&prm, 0, sizeof(prm));
    const int in = d & 1;
    if ((rc = make_map(&prm.tmA_hi, h_hi[in], w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmA_lo, h_lo[in], w.rows_pad, H, w.kp_h, TM))) return rc;
    if ((rc = make_map(&prm.tmB_hi, w_hi(d + 1), H, H, wl.ld[d + 1], BN))) return rc;
    if ((rc = make_map(&prm.tmB_lo, w_lo(d + 1), H, H, wl.ld[d + 1], BN))) return rc;
    prm.num_k = (int)cgr_ceil_div(H, BK);
    prm.n_total = H;
    prm.unscale = unscale + d + 1;
    prm.tile_info = g->tile_info;
    prm.in_ptr = g->in_ptr; prm.in_idx = g->in_idx; prm.src = g->src; prm.atom_ptr = g->atom_ptr;
    prm.act = p->act;
    prm.Q = PQ + H; prm.ldq = 2 * H;
    prm.w_ffn = p->w_ffn;
    prm.partial_out = partial;
    prm.n_rxn = B;
    prm.overflow = flag;
    rc = launch_gemm<EPI_READOUT>(prm, (int)T, w.n_slices, "tc_readout", st);
    if (rc) return rc;
  }
  {
    CgrRange prof("tc_finalize", st);
    cgr_note_launch("tc_finalize", st, 1);
    tc_finalize_kernel<<<(unsigned)cgr_ceil_div(B, 256), 256, 0, st>>>(partial, w.n_slices, B, p->b_ffn, out);
    CGR_LAUNCH_CHECK();
  }
  if (g->tc_status) {      // sticky overflow report for the caller (checked lazily on the host)
    CGR_CUDA(cudaMemcpyAsync(g->tc_status, flag, sizeof(int), cudaMemcpyDeviceToDevice, st));
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
  split_rows_kernel<<<(unsigned)cgr_ceil_div(M, 8), 256, 0, st>>>(x, ldx, M, (int)K, a_hi, a_lo, kp, flag);
  split_rows_kernel<<<(unsigned)cgr_ceil_div(N, 8), 256, 0, st>>>(wgt, ldw, N, (int)K, b_hi, b_lo, kp, flag);
  CGR_LAUNCH_CHECK();
  TcGemmParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  if ((rc = make_map(&prm.tmA_hi, a_hi, M, K, kp, TM))) return rc;
  if ((rc = make_map(&prm.tmA_lo, a_lo, M, K, kp, TM))) return rc;
  if ((rc = make_map(&prm.tmB_hi, b_hi, N, K, kp, BN))) return rc;
  if ((rc = make_map(&prm.tmB_lo, b_lo, N, K, kp, BN))) return rc;
  prm.num_k = (int)cgr_ceil_div(K, BK);
  prm.n_total = (int)N;
  prm.m_rows = (int)M;
  prm.unscale = one;
  prm.bias = bias;
  prm.out_f32 = out;
  prm.ldc = N;
  return launch_gemm<EPI_PLAIN>(prm, (int)cgr_ceil_div(M, TM), (int)cgr_ceil_div(N, BN), "tc_linear", st);
}
