// Batch collation and CSR (a2b / b2a / b2revb) construction — integer work, bit-exact.
//
// Replaces the host-side PyG collate the reference loaders run
// (cgr_mpnn_3D/training/trainer.py:105-118, test.py:85-90) and derives the index arrays the
// message passing needs from the batched edge_index (cgr_mpnn_3D/models/GNN.py:85,132-138).
#include "common.cuh"
#include "../../include/cgr_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// exclusive scan of one value per thread across a 256-thread block; returns block total via *total
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total) {
  __shared__ T warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      T t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;  // exclusive warp offsets
    if (lane == SCAN_THREADS / 32 - 1) *total = wi;
  }
  __syncthreads();
  T res = warp_sums[warp] + incl - v;
  __syncthreads();
  return res;
}

// phase A: per-tile totals
template <typename TIn, typename TOut>
__global__ void scan_tile_sums(const TIn* __restrict__ in, int64_t n, TOut* __restrict__ tile_sums) {
  __shared__ TOut total;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  TOut s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += (TOut)in[base + i];
  block_exclusive_scan<TOut>(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// phase B: one block turns tile totals into exclusive tile offsets (any count, carried chunks)
template <typename TOut>
__global__ void scan_tile_offsets(TOut* __restrict__ tile_sums, int64_t n_tiles) {
  __shared__ TOut total;
  __shared__ TOut carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    const int64_t i = base + threadIdx.x;
    TOut v = i < n_tiles ? tile_sums[i] : TOut(0);
    TOut ex = block_exclusive_scan<TOut>(v, &total);
    TOut carry = carry_s;
    if (i < n_tiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
}

// phase C: exclusive scan written to out[0..n], out[n] = grand total
template <typename TIn, typename TOut>
__global__ void scan_write(const TIn* __restrict__ in, int64_t n, const TOut* __restrict__ tile_offsets,
                           TOut* __restrict__ out) {
  __shared__ TOut total;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  TOut v[SCAN_ITEMS];
  TOut s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? (TOut)in[base + i] : TOut(0);
    s += v[i];
  }
  TOut ex = block_exclusive_scan<TOut>(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
    if (base + i == n - 1) out[n] = ex;
  }
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

template <typename TIn, typename TOut>
int exclusive_scan(const TIn* in, int64_t n, TOut* out, TOut* tile_ws, cudaStream_t st) {
  const int64_t n_tiles = n > 0 ? cgr_ceil_div(n, SCAN_TILE) : 1;
  scan_tile_sums<TIn, TOut><<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(in, n, tile_ws);
  scan_tile_offsets<TOut><<<1, SCAN_THREADS, 0, st>>>(tile_ws, n_tiles);
  scan_write<TIn, TOut><<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(in, n, tile_ws, out);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

// largest g with ptr[g] <= i  (ptr non-decreasing, ptr[0] == 0, ptr[nb] == total > i)
__device__ __forceinline__ int64_t segment_of(const int64_t* __restrict__ ptr, int64_t nb, int64_t i) {
  int64_t lo = 0, hi = nb;  // invariant: ptr[lo] <= i < ptr[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void collate_edges_kernel(const int64_t* __restrict__ local_ei, const int64_t* __restrict__ ptr,
                                     const int64_t* __restrict__ edge_ptr, int64_t nb, int64_t ne,
                                     int64_t* __restrict__ edge_index) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int64_t off = ptr[segment_of(edge_ptr, nb, e)];
  edge_index[e] = local_ei[e] + off;
  edge_index[ne + e] = local_ei[ne + e] + off;
}

__global__ void collate_batch_kernel(const int64_t* __restrict__ ptr, int64_t nb, int64_t na,
                                     int64_t* __restrict__ batch) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= na) return;
  batch[v] = segment_of(ptr, nb, v);
}

// ---- CSR -------------------------------------------------------------------------------------

__global__ void csr_count_kernel(const int64_t* __restrict__ ei, int64_t ne, int64_t na,
                                 int32_t* __restrict__ src, int32_t* __restrict__ dst,
                                 int32_t* __restrict__ deg, int32_t* __restrict__ status) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int64_t s = ei[e], d = ei[ne + e];
  int flag = 0;
  if (s < 0 || s >= na || d < 0 || d >= na) flag |= 2;
  const int64_t r = e ^ 1;                       // b2revb: reference GNN.py:136-138
  if (r >= ne || ei[r] != d || ei[ne + r] != s) flag |= 1;
  src[e] = (int32_t)s;
  dst[e] = (int32_t)d;
  if (!(flag & 2)) atomicAdd(&deg[d], 1);
  if (flag) atomicOr(status, flag);
}

__global__ void csr_fill_kernel(const int32_t* __restrict__ dst, int64_t ne, int64_t na,
                                const int32_t* __restrict__ in_ptr, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ in_idx) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int32_t d = dst[e];
  if (d < 0 || d >= na) return;
  const int32_t pos = in_ptr[d] + atomicAdd(&cursor[d], 1);
  in_idx[pos] = (int32_t)e;
}

// ascending bond id inside each atom's group == the order CPU scatter_add_ accumulates in
__global__ void csr_sort_kernel(const int32_t* __restrict__ in_ptr, int64_t na, int32_t* __restrict__ in_idx,
                                int32_t* __restrict__ status) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= na) return;
  const int32_t b = in_ptr[v], e = in_ptr[v + 1];
  if (e == b) atomicOr(status, 4);
  for (int32_t i = b + 1; i < e; ++i) {
    const int32_t key = in_idx[i];
    int32_t j = i - 1;
    while (j >= b && in_idx[j] > key) {
      in_idx[j + 1] = in_idx[j];
      --j;
    }
    in_idx[j + 1] = key;
  }
}

__global__ void atom_ptr_kernel(const int64_t* __restrict__ batch, int64_t na, int64_t nb,
                                int32_t* __restrict__ atom_ptr) {
  // atom_ptr[g] = first v with batch[v] >= g; batch sorted ascending (PyG collate)
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > na) return;
  const int64_t cur = v < na ? batch[v] : nb;
  const int64_t prev = v > 0 ? batch[v - 1] : -1;
  for (int64_t g = prev + 1; g <= cur && g <= nb; ++g) atom_ptr[g] = (int32_t)v;
}

// One block per reaction: the whole CSR of a collated batch in ONE launch when per-reaction offsets are known.
// Same result as cgr_csr_build (bit-exact), used by the host-buffer inference entry.
constexpr int RX_MAX = 256;     // atoms / bonds of one reaction handled in shared memory
__global__ void __launch_bounds__(128) csr_by_reaction_kernel(const int64_t* __restrict__ ei, const int32_t* __restrict__ edge_ptr,
                                                              const int32_t* __restrict__ atom_ptr, int64_t ne, int64_t na,
                                                              int64_t nb, int32_t* __restrict__ src, int32_t* __restrict__ dst,
                                                              int32_t* __restrict__ in_ptr, int32_t* __restrict__ in_idx,
                                                              int32_t* __restrict__ status,
                                                              const int32_t* __restrict__ rxn_shift) {
  __shared__ int deg[RX_MAX], off[RX_MAX + 1], cur[RX_MAX];
  __shared__ int32_t loc[RX_MAX];
  const int g = blockIdx.x;
  const int e0 = edge_ptr[g], e1 = edge_ptr[g + 1], a0 = atom_ptr[g], a1 = atom_ptr[g + 1];
  const int n_e = e1 - e0, n_a = a1 - a0;
  if (n_e > RX_MAX || n_a > RX_MAX || n_e < 0 || n_a < 0) { if (threadIdx.x == 0) atomicOr(status, 8); return; }
  for (int v = threadIdx.x; v < n_a; v += blockDim.x) { deg[v] = 0; cur[v] = 0; }
  __syncthreads();
  int flag = 0;
  const int64_t sh = rxn_shift ? rxn_shift[g] : 0;     // atom-id shift of the host batch this reaction came from
  for (int j = threadIdx.x; j < n_e; j += blockDim.x) {
    const int64_t e = e0 + j, s = ei[e] + sh, d = ei[ne + e] + sh;
    if (s < a0 || s >= a1 || d < a0 || d >= a1) flag |= 2;
    const int64_t r = e ^ 1;
    if (r >= ne || ei[r] + sh != d || ei[ne + r] + sh != s) flag |= 1;
    src[e] = (int32_t)s;
    dst[e] = (int32_t)d;
    if (!(flag & 2)) atomicAdd(&deg[d - a0], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int v = 0; v < n_a; ++v) { off[v] = acc; acc += deg[v]; }
    off[n_a] = acc;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < n_a; v += blockDim.x) {
    in_ptr[a0 + v] = e0 + off[v];
    if (deg[v] == 0) flag |= 4;
  }
  if (g == nb - 1 && threadIdx.x == 0) in_ptr[na] = (int32_t)ne;
  for (int j = threadIdx.x; j < n_e; j += blockDim.x) {
    const int d = dst[e0 + j] - a0;
    if (d >= 0 && d < n_a) loc[off[d] + atomicAdd(&cur[d], 1)] = e0 + j;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < n_a; v += blockDim.x) {          // ascending bond id inside each atom's group
    const int b = off[v], e = off[v + 1];
    for (int i = b + 1; i < e; ++i) {
      const int32_t key = loc[i];
      int k = i - 1;
      while (k >= b && loc[k] > key) { loc[k + 1] = loc[k]; --k; }
      loc[k + 1] = key;
    }
    for (int i = b; i < e; ++i) in_idx[e0 + i] = loc[i];
  }
  if (flag) atomicOr(status, flag);
}

}  // namespace

int csr_by_reaction_shifted(const int64_t* edge_index, const int32_t* edge_ptr, const int32_t* atom_ptr,
                            const int32_t* rxn_shift, int64_t n_rxn, int64_t n_bonds, int64_t n_atoms, int32_t* src,
                            int32_t* dst, int32_t* in_ptr, int32_t* in_idx, int32_t* status, cudaStream_t st) {
  CGR_CHECK_ARG(edge_index && edge_ptr && atom_ptr && src && dst && in_ptr && in_idx && status && n_rxn > 0,
                "cgr_csr_build_by_reaction: bad argument");
  cgr_note_launch("csr_by_reaction", st, 1);
  csr_by_reaction_kernel<<<(unsigned)n_rxn, 128, 0, st>>>(edge_index, edge_ptr, atom_ptr, n_bonds, n_atoms, n_rxn, src,
                                                          dst, in_ptr, in_idx, status, rxn_shift);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

extern "C" int cgr_csr_build_by_reaction(const int64_t* edge_index, const int32_t* edge_ptr, const int32_t* atom_ptr,
                                         int64_t n_rxn, int64_t n_bonds, int64_t n_atoms, int32_t* src, int32_t* dst,
                                         int32_t* in_ptr, int32_t* in_idx, int32_t* status, void* stream) {
  return csr_by_reaction_shifted(edge_index, edge_ptr, atom_ptr, nullptr, n_rxn, n_bonds, n_atoms, src, dst, in_ptr,
                                 in_idx, status, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// Device-resident reaction store: the whole featurised data set lives packed in HBM
// (per-reaction rows are contiguous) and a batch is ASSEMBLED ON THE DEVICE from a list of reaction ids -- what
// ChemDataset.__getitem__ (data/ChemDataset.py:69-94) + PyG collate (trainer.py:105-118) do per item on the host.
// SG_PARTS blocks per selected reaction: vector copies of its atom / bond rows, edge_index shifted to batch-global ids,
// batch vector, label.  HBM-bound copy; outputs are bit-identical to collating the same reactions on the host.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int SG_PARTS = 8;          // blocks per selected reaction (a T1x reaction moves ~60 kB)

// contiguous copy of `total` floats by part `part` of `parts` blocks, widest vector both pointers allow
__device__ __forceinline__ void block_copy(const float* __restrict__ src, float* __restrict__ dst, int64_t total, int part,
                                           int parts) {
  const uintptr_t mis = (uintptr_t)src | (uintptr_t)dst;
  const int64_t tid = (int64_t)part * blockDim.x + threadIdx.x, nth = (int64_t)parts * blockDim.x;
  if ((mis & 15) == 0) {
    const int64_t v = total >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
    for (int64_t i = tid; i < v; i += nth) d4[i] = __ldg(s4 + i);
    for (int64_t i = (v << 2) + tid; i < total; i += nth) dst[i] = __ldg(src + i);
  } else if ((mis & 7) == 0) {
    const int64_t v = total >> 1;
    const float2* s2 = reinterpret_cast<const float2*>(src);
    float2* d2 = reinterpret_cast<float2*>(dst);
#pragma unroll 4
    for (int64_t i = tid; i < v; i += nth) d2[i] = __ldg(s2 + i);
    for (int64_t i = (v << 1) + tid; i < total; i += nth) dst[i] = __ldg(src + i);
  } else {
#pragma unroll 4
    for (int64_t i = tid; i < total; i += nth) dst[i] = __ldg(src + i);
  }
}

__global__ void __launch_bounds__(256) store_gather_kernel(
    const float* __restrict__ x_all, const float* __restrict__ ea_all, const int32_t* __restrict__ ei_all,
    const int64_t* __restrict__ node_ptr, const int64_t* __restrict__ edge_ptr, const float* __restrict__ y_all,
    int64_t e_all, const int64_t* __restrict__ sel, const int64_t* __restrict__ out_node_ptr,
    const int64_t* __restrict__ out_edge_ptr, int fa, int fb, int64_t e_out, float* __restrict__ x,
    float* __restrict__ ea, int64_t* __restrict__ ei, int64_t* __restrict__ batch, float* __restrict__ y) {
  const int64_t b = blockIdx.y;
  const int part = blockIdx.x;
  const int64_t r = sel[b];
  const int64_t a_in = node_ptr[r], n = node_ptr[r + 1] - a_in, a_out = out_node_ptr[b];
  const int64_t e_in = edge_ptr[r], e = edge_ptr[r + 1] - e_in, e_o = out_edge_ptr[b];
  block_copy(x_all + a_in * fa, x + a_out * fa, n * fa, part, SG_PARTS);                    // atom rows
  if (fb > 0) block_copy(ea_all + e_in * fb, ea + e_o * fb, e * fb, part, SG_PARTS);         // bond rows
  if (part == 0) {
    for (int64_t j = threadIdx.x; j < e; j += blockDim.x) {      // local atom ids -> batch-global ids (int64 like PyG)
      ei[e_o + j] = (int64_t)__ldg(ei_all + e_in + j) + a_out;
      ei[e_out + e_o + j] = (int64_t)__ldg(ei_all + e_all + e_in + j) + a_out;
    }
  } else if (part == 1) {
    if (batch)
      for (int64_t v = threadIdx.x; v < n; v += blockDim.x) batch[a_out + v] = b;
    if (y && threadIdx.x == 0) y[b] = __ldg(y_all + r);
  }
}

// The screening loop (cgr_store_infer) never needs the batch's fp32 atom features: the tcgen05 forward reads them as the
// FP16 (hi, lo) operand pair of the atom projection.  This variant of the gather writes that pair straight from the
// store rows -- the same arithmetic as split_rows_kernel (tc.cu), so the operands are bit-identical -- instead of an
// fp32 copy that a second kernel would read back: 8 bytes per feature over HBM instead of 16.
// Work of a reaction is flattened over its SG_PARTS blocks in pairs of columns (fa even) or single columns.
template <bool PAIRS>
__global__ void __launch_bounds__(256) store_gather_split_kernel(
    const float* __restrict__ x_all, const float* __restrict__ ea_all, const int32_t* __restrict__ ei_all,
    const int64_t* __restrict__ node_ptr, const int64_t* __restrict__ edge_ptr, int64_t e_all,
    const int64_t* __restrict__ sel, const int64_t* __restrict__ out_node_ptr, const int64_t* __restrict__ out_edge_ptr,
    int fa, int fb, int64_t e_out, __half* __restrict__ x_hi, __half* __restrict__ x_lo, int64_t ldo,
    float* __restrict__ ea, int64_t* __restrict__ ei, int* __restrict__ range_flag, int flag_bit) {
  const int64_t b = blockIdx.y;
  const int part = blockIdx.x;
  const int64_t r = sel[b];
  const int64_t a_in = node_ptr[r], n = node_ptr[r + 1] - a_in, a_out = out_node_ptr[b];
  const int64_t e_in = edge_ptr[r], e = edge_ptr[r + 1] - e_in, e_o = out_edge_ptr[b];
  const float* __restrict__ xs = x_all + a_in * fa;
  __half* __restrict__ hi = x_hi + a_out * ldo;
  __half* __restrict__ lo = x_lo + a_out * ldo;
  const int tid = part * blockDim.x + threadIdx.x, nth = SG_PARTS * blockDim.x;
  bool ovf = false;
  if (PAIRS) {                                           // rows start 8-byte aligned (fa even), outputs 4-byte (ldo even)
    const int w2 = fa >> 1, total = (int)n * w2;
    const float2* __restrict__ s2 = reinterpret_cast<const float2*>(xs);
    const int dq = nth / w2, dr = nth - dq * w2;         // (row, column pair) advance by nth without a division per item
    int row = tid / w2, c2 = tid - row * w2;
#pragma unroll 4
    for (int i = tid; i < total; i += nth, row += dq, c2 += dr) {
      if (c2 >= w2) { c2 -= w2; ++row; }
      const float2 v = __ldg(s2 + i);
      ovf |= fabsf(v.x) > 60000.f || fabsf(v.y) > 60000.f;
      const __half h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y);
      const __half l0 = __float2half_rn(v.x - __half2float(h0)), l1 = __float2half_rn(v.y - __half2float(h1));
      const int64_t o = (int64_t)row * ldo + 2 * c2;
      *reinterpret_cast<__half2*>(hi + o) = __halves2half2(h0, h1);
      *reinterpret_cast<__half2*>(lo + o) = __halves2half2(l0, l1);
    }
  } else {
    const int total = (int)n * fa;
    const int dq = nth / fa, dr = nth - dq * fa;
    int row = tid / fa, c = tid - row * fa;
#pragma unroll 4
    for (int i = tid; i < total; i += nth, row += dq, c += dr) {
      if (c >= fa) { c -= fa; ++row; }
      const float v = __ldg(xs + i);
      ovf |= fabsf(v) > 60000.f;
      const __half h = __float2half_rn(v);
      hi[(int64_t)row * ldo + c] = h;
      lo[(int64_t)row * ldo + c] = __float2half_rn(v - __half2float(h));
    }
  }
  if (ovf) atomicOr(range_flag, flag_bit);
  if (fb > 0) block_copy(ea_all + e_in * fb, ea + e_o * fb, e * fb, part, SG_PARTS);         // bond rows
  if (part == 0) {
    for (int64_t j = threadIdx.x; j < e; j += blockDim.x) {      // local atom ids -> batch-global ids (int64 like PyG)
      ei[e_o + j] = (int64_t)__ldg(ei_all + e_in + j) + a_out;
      ei[e_out + e_o + j] = (int64_t)__ldg(ei_all + e_all + e_in + j) + a_out;
    }
  }
}
}  // namespace

// internal (api.cu: cgr_store_infer): assemble the batch `sel` with the atom features as the FP16 (hi, lo) operand pair
int store_gather_split(const float* x_all, const float* ea_all, const int32_t* ei_all, const int64_t* node_ptr,
                       const int64_t* edge_ptr, int64_t e_all, const int64_t* sel, const int64_t* out_node_ptr,
                       const int64_t* out_edge_ptr, int64_t n_sel, int32_t fa, int32_t fb, int64_t e_out, void* x_hi,
                       void* x_lo, int64_t ldo, float* edge_attr, int64_t* edge_index, int* range_flag, int flag_bit,
                       cudaStream_t st) {
  CGR_CHECK_ARG(x_all && ei_all && node_ptr && edge_ptr && sel && out_node_ptr && out_edge_ptr && x_hi && x_lo &&
                    edge_index && range_flag && (fb == 0 || (ea_all && edge_attr)),
                "store_gather_split: null pointer");
  CGR_CHECK_ARG(n_sel >= 0 && fa > 0 && fb >= 0 && e_all >= 0 && e_out >= 0 && ldo >= fa, "store_gather_split: bad size");
  if (n_sel == 0) return CGR_OK;
  cgr_note_launch("store_gather_split", st, 1);
  const dim3 grid(SG_PARTS, (unsigned)n_sel);
  if ((fa & 1) == 0 && (ldo & 1) == 0 && ((uintptr_t)x_all & 7) == 0 && ((uintptr_t)x_hi & 3) == 0 && ((uintptr_t)x_lo & 3) == 0)
    store_gather_split_kernel<true><<<grid, 256, 0, st>>>(x_all, ea_all, ei_all, node_ptr, edge_ptr, e_all, sel, out_node_ptr,
                                                           out_edge_ptr, fa, fb, e_out, (__half*)x_hi, (__half*)x_lo, ldo,
                                                           edge_attr, edge_index, range_flag, flag_bit);
  else
    store_gather_split_kernel<false><<<grid, 256, 0, st>>>(x_all, ea_all, ei_all, node_ptr, edge_ptr, e_all, sel, out_node_ptr,
                                                            out_edge_ptr, fa, fb, e_out, (__half*)x_hi, (__half*)x_lo, ldo,
                                                            edge_attr, edge_index, range_flag, flag_bit);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

extern "C" int cgr_store_gather(const float* x_all, const float* ea_all, const int32_t* ei_all, const int64_t* node_ptr,
                                const int64_t* edge_ptr, const float* y_all, int64_t e_all, const int64_t* sel,
                                const int64_t* out_node_ptr, const int64_t* out_edge_ptr, int64_t n_sel, int32_t fa,
                                int32_t fb, int64_t e_out, float* x, float* edge_attr, int64_t* edge_index,
                                int64_t* batch, float* y, void* stream) {
  CGR_CHECK_ARG(x_all && ei_all && node_ptr && edge_ptr && sel && out_node_ptr && out_edge_ptr && (y_all || !y),
                "cgr_store_gather: null input");
  CGR_CHECK_ARG(fb == 0 || ea_all, "cgr_store_gather: edge_attr store missing");
  CGR_CHECK_ARG(x && edge_index && (fb == 0 || edge_attr), "cgr_store_gather: null output");   // batch / y optional
  CGR_CHECK_ARG(n_sel >= 0 && fa > 0 && fb >= 0 && e_all >= 0 && e_out >= 0, "cgr_store_gather: bad size");
  if (n_sel == 0) return CGR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cgr_note_launch("store_gather", st, 1);
  store_gather_kernel<<<dim3(SG_PARTS, (unsigned)n_sel), 256, 0, st>>>(x_all, ea_all, ei_all, node_ptr, edge_ptr, y_all, e_all, sel,
                                                       out_node_ptr, out_edge_ptr, fa, fb, e_out, x, edge_attr,
                                                       edge_index, batch, y);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

extern "C" size_t cgr_collate_workspace(int64_t n_rxn) {
  return (size_t)(cgr_ceil_div(n_rxn > 0 ? n_rxn : 1, SCAN_TILE) + 1) * sizeof(int64_t);
}

extern "C" int cgr_collate_indices(const int64_t* n_nodes, const int64_t* n_edges,
                                   const int64_t* local_edge_index, int64_t n_rxn, int64_t n_bonds,
                                   int64_t n_atoms, int64_t* edge_index, int64_t* batch, int64_t* ptr,
                                   int64_t* edge_ptr, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  CGR_CHECK_ARG(n_rxn >= 0 && n_bonds >= 0 && n_atoms >= 0, "cgr_collate_indices: negative size");
  CGR_CHECK_ARG(workspace_bytes >= cgr_collate_workspace(n_rxn), "cgr_collate_indices: workspace too small");
  CGR_CHECK_ARG(ptr && edge_ptr && workspace, "cgr_collate_indices: null output");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t* tile_ws = (int64_t*)workspace;
  int rc = exclusive_scan<int64_t, int64_t>(n_nodes, n_rxn, ptr, tile_ws, st);
  if (rc) return rc;
  rc = exclusive_scan<int64_t, int64_t>(n_edges, n_rxn, edge_ptr, tile_ws, st);
  if (rc) return rc;
  if (n_bonds > 0) {
    collate_edges_kernel<<<(unsigned)cgr_ceil_div(n_bonds, 256), 256, 0, st>>>(
        local_edge_index, ptr, edge_ptr, n_rxn, n_bonds, edge_index);
  }
  if (n_atoms > 0) {
    collate_batch_kernel<<<(unsigned)cgr_ceil_div(n_atoms, 256), 256, 0, st>>>(ptr, n_rxn, n_atoms, batch);
  }
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

extern "C" size_t cgr_csr_workspace(int64_t n_atoms, int64_t n_bonds) {
  (void)n_bonds;
  const size_t na = (size_t)(n_atoms > 0 ? n_atoms : 1);
  return cgr_align_up(2 * na * sizeof(int32_t), 256) +
         (size_t)(cgr_ceil_div((int64_t)na, SCAN_TILE) + 1) * sizeof(int32_t);
}

extern "C" int cgr_csr_build(const int64_t* edge_index, int64_t n_bonds, int64_t n_atoms, int32_t* src,
                             int32_t* dst, int32_t* in_ptr, int32_t* in_idx, int32_t* status,
                             void* workspace, size_t workspace_bytes, void* stream) {
  CGR_CHECK_ARG(n_bonds >= 0 && n_atoms >= 0, "cgr_csr_build: negative size");
  CGR_CHECK_ARG(n_bonds < (1ll << 31) && n_atoms < (1ll << 31), "cgr_csr_build: sizes exceed int32 index range");
  CGR_CHECK_ARG(workspace_bytes >= cgr_csr_workspace(n_atoms, n_bonds), "cgr_csr_build: workspace too small");
  CGR_CHECK_ARG(src && dst && in_ptr && in_idx && status && workspace, "cgr_csr_build: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t na = (size_t)(n_atoms > 0 ? n_atoms : 1);
  int32_t* deg = (int32_t*)workspace;
  int32_t* cursor = deg + na;
  int32_t* tile_ws = (int32_t*)((char*)workspace + cgr_align_up(2 * na * sizeof(int32_t), 256));
  CGR_CUDA(cudaMemsetAsync(deg, 0, 2 * na * sizeof(int32_t), st));
  CGR_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  if (n_bonds > 0)
    csr_count_kernel<<<(unsigned)cgr_ceil_div(n_bonds, 256), 256, 0, st>>>(edge_index, n_bonds, n_atoms, src,
                                                                           dst, deg, status);
  int rc = exclusive_scan<int32_t, int32_t>(deg, n_atoms, in_ptr, tile_ws, st);
  if (rc) return rc;
  if (n_bonds > 0)
    csr_fill_kernel<<<(unsigned)cgr_ceil_div(n_bonds, 256), 256, 0, st>>>(dst, n_bonds, n_atoms, in_ptr,
                                                                          cursor, in_idx);
  if (n_atoms > 0)
    csr_sort_kernel<<<(unsigned)cgr_ceil_div(n_atoms, 128), 128, 0, st>>>(in_ptr, n_atoms, in_idx, status);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

extern "C" int cgr_atom_ptr_from_batch(const int64_t* batch, int64_t n_atoms, int64_t n_rxn,
                                       int32_t* atom_ptr, void* stream) {
  CGR_CHECK_ARG(n_atoms >= 0 && n_rxn >= 0 && atom_ptr, "cgr_atom_ptr_from_batch: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  atom_ptr_kernel<<<(unsigned)cgr_ceil_div(n_atoms + 1, 256), 256, 0, st>>>(batch, n_atoms, n_rxn, atom_ptr);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}
