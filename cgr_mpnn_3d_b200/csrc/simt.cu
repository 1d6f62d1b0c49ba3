// SIMT fp32 engine (CGR_ENGINE_SIMT): exact-fp32 FMA GEMM + CSR gathers + deterministic reductions.
// Layer-wise formulation of cgr_mpnn_3D/models/GNN.py:76-145; no atomics on floating data, so every
// result is run-to-run bit-stable (the reference's CUDA path uses fp32 atomics).
#include "simt.cuh"

namespace {

constexpr int GBM = 64, GBN = 64, GBK = 16, GPAD = 4;

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int64_t M, N, K;
  int64_t k_chunk;       // K range per blockIdx.z (multiple of GBK)
  float* partial;        // non-null => split-K partial output [z][M][N]
  GemmEpilogue epi;
};

template <bool AK, bool BK>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[GBK][GBN + GPAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * GBM, n0 = (int64_t)blockIdx.x * GBN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  const int64_t kend = kbeg + g.k_chunk < g.K ? kbeg + g.k_chunk : g.K;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int64_t k0) {
    if (AK) {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      const int64_t m = m0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t k = k0 + kq + j;
        ra[j] = (m < g.M && k < kend) ? __ldg(g.A + m * g.lda + k) : 0.f;
      }
    } else {
      const int kk = tid >> 4, mq = (tid & 15) * 4;
      const int64_t k = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t m = m0 + mq + j;
        ra[j] = (m < g.M && k < kend) ? __ldg(g.A + k * g.lda + m) : 0.f;
      }
    }
    if (BK) {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      const int64_t n = n0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t k = k0 + kq + j;
        rb[j] = (n < g.N && k < kend) ? __ldg(g.B + n * g.ldb + k) : 0.f;
      }
    } else {
      const int kk = tid >> 4, nq = (tid & 15) * 4;
      const int64_t k = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + nq + j;
        rb[j] = (n < g.N && k < kend) ? __ldg(g.B + k * g.ldb + n) : 0.f;
      }
    }
  };
  auto store_tiles = [&]() {
    if (AK) {
      const int r = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) As[kq + j][r] = ra[j];
    } else {
      const int kk = tid >> 4, mq = (tid & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) As[kk][mq + j] = ra[j];
    }
    if (BK) {
      const int r = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[kq + j][r] = rb[j];
    } else {
      const int kk = tid >> 4, nq = (tid & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[kk][nq + j] = rb[j];
    }
  };

  if (kbeg < kend) load_tiles(kbeg);
  for (int64_t k0 = kbeg; k0 < kend; k0 += GBK) {
    store_tiles();
    __syncthreads();
    if (k0 + GBK < kend) load_tiles(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (g.partial) {
    float* P = g.partial + (int64_t)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx * 4 + j;
        if (n < g.N) P[m * g.N + n] = acc[i][j];
      }
    }
    return;
  }

  const GemmEpilogue& e = g.epi;
  const float rs = e.res ? (e.res_scale ? __ldg(e.res_scale) : 1.f) : 0.f;
  const float keep_scale = e.dropout_p > 0.f ? 1.f / (1.f - e.dropout_p) : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (e.bias) v += __ldg(e.bias + n);
      if (e.res) v = fmaf(rs, __ldg(e.res + m * e.ldr + n), v);
      if (e.preact) e.preact[m * g.ldc + n] = v;
      v = cgr_act(v, e.act);
      if (e.dropout_p > 0.f)
        v = cgr_dropout_keep(e.seed, e.layer, (uint64_t)(m * g.N + n), e.dropout_p) ? v * keep_scale : 0.f;
      g.C[m * g.ldc + n] = v;
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t M, int64_t N,
                                     float* __restrict__ C, int64_t ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * M * N + i];   // fixed order
  C[(i / N) * ldc + (i % N)] = s;
}

// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float fprime_at(const GatherPost& p, int64_t idx) {
  float sc = 1.f;
  if (p.dropout_p > 0.f) {
    sc = 1.f / (1.f - p.dropout_p);
    if (p.act != CGR_ACT_RELU && !cgr_dropout_keep(p.seed, p.layer, (uint64_t)idx, p.dropout_p)) return 0.f;
  }
  if (p.act == CGR_ACT_RELU) return __ldg(p.h_next + idx) > 0.f ? sc : 0.f;
  return sc * cgr_act_grad(__ldg(p.z + idx), 0.f, p.act);
}

// one warp per bond; lanes sweep the feature row (float4 when H % 4 == 0)
template <int VEC>
__global__ void __launch_bounds__(256) gather_bonds_kernel(const float* __restrict__ in, const int32_t* __restrict__ node,
                                                           const int32_t* __restrict__ in_ptr,
                                                           const int32_t* __restrict__ in_idx, int flip,
                                                           float* __restrict__ out, int64_t E, int H,
                                                           const GatherPost post) {
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= E) return;
  const int lane = threadIdx.x & 31;
  const int32_t v = __ldg(node + e);
  const int32_t beg = __ldg(in_ptr + v), end = __ldg(in_ptr + v + 1);
  const int64_t rev = e ^ 1;
  // neighbour ids fetched once by the lanes and broadcast, instead of a dependent load chain per column group
  const int deg = end - beg;
  const int32_t my_nb = lane < deg ? (__ldg(in_idx + beg + lane) ^ flip) : 0;
  int32_t nbr[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) nbr[p] = __shfl_sync(0xffffffffu, my_nb, p);   // every lane takes part, before any divergence
  if (VEC == 4) {
    const int H4 = H >> 2;
    for (int c = lane; c < H4; c += 32) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < 8; ++p) {                  // ascending bond id: CPU scatter_add_ order
        if (p < deg) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(in + (int64_t)nbr[p] * H) + c);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
      }
      for (int32_t p = 8; p < deg; ++p) {
        const int64_t k = (int64_t)(__ldg(in_idx + beg + p) ^ flip);
        const float4 t = __ldg(reinterpret_cast<const float4*>(in + k * H) + c);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      const float4 r = __ldg(reinterpret_cast<const float4*>(in + rev * H) + c);
      a.x -= r.x; a.y -= r.y; a.z -= r.z; a.w -= r.w;
      const int64_t idx = e * H + 4 * c;
      if (post.mode == 2) {
        const float4 ad = __ldg(reinterpret_cast<const float4*>(post.add + idx));
        a.x += ad.x; a.y += ad.y; a.z += ad.z; a.w += ad.w;
      }
      if (post.mode) {
        a.x *= fprime_at(post, idx); a.y *= fprime_at(post, idx + 1);
        a.z *= fprime_at(post, idx + 2); a.w *= fprime_at(post, idx + 3);
      }
      reinterpret_cast<float4*>(out + e * H)[c] = a;
    }
  } else {
    for (int c = lane; c < H; c += 32) {
      float a = 0.f;
      for (int32_t p = beg; p < end; ++p) a += __ldg(in + (int64_t)(__ldg(in_idx + p) ^ flip) * H + c);
      a -= __ldg(in + rev * H + c);
      const int64_t idx = e * H + c;
      if (post.mode == 2) a += __ldg(post.add + idx);
      if (post.mode) a *= fprime_at(post, idx);
      out[idx] = a;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) atom_sum_kernel(const float* __restrict__ in, const int32_t* __restrict__ in_ptr,
                                                       const int32_t* __restrict__ in_idx, int flip,
                                                       float* __restrict__ out, int64_t N, int H) {
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= N) return;
  const int lane = threadIdx.x & 31;
  const int32_t beg = __ldg(in_ptr + v), end = __ldg(in_ptr + v + 1);
  if (VEC == 4) {
    const int H4 = H >> 2;
    for (int c = lane; c < H4; c += 32) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int32_t p = beg; p < end; ++p) {
        const int64_t k = (int64_t)(__ldg(in_idx + p) ^ flip);
        const float4 t = __ldg(reinterpret_cast<const float4*>(in + k * H) + c);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      reinterpret_cast<float4*>(out + v * H)[c] = a;
    }
  } else {
    for (int c = lane; c < H; c += 32) {
      float a = 0.f;
      for (int32_t p = beg; p < end; ++p) a += __ldg(in + (int64_t)(__ldg(in_idx + p) ^ flip) * H + c);
      out[v * H + c] = a;
    }
  }
}

__global__ void __launch_bounds__(256) expand_dst_kernel(const float* __restrict__ ds, const int32_t* __restrict__ dst,
                                                         float* __restrict__ dz, int64_t E, int H,
                                                         const GatherPost post) {
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= E) return;
  const int lane = threadIdx.x & 31;
  const int64_t v = dst ? (int64_t)__ldg(dst + e) : e;       // dst == null: row-wise (dz = ds . fprime)
  for (int c = lane; c < H; c += 32) {
    const int64_t idx = e * H + c;
    dz[idx] = __ldg(ds + v * H + c) * (post.mode ? fprime_at(post, idx) : 1.f);
  }
}

__global__ void __launch_bounds__(128) edge_init_kernel(const float* __restrict__ P, const float* __restrict__ ea,
                                                        const int32_t* __restrict__ src,
                                                        const float* __restrict__ w_init,
                                                        const float* __restrict__ b_init, int64_t E, int fa,
                                                        int fb, int H, int act, float* __restrict__ h0,
                                                        float* __restrict__ z0) {
  extern __shared__ float ea_s[];                 // [4][fb]
  const int64_t e0 = (int64_t)blockIdx.x * 4;
  for (int i = threadIdx.x; i < 4 * fb; i += blockDim.x) {
    const int64_t e = e0 + i / fb;
    ea_s[i] = e < E ? __ldg(ea + e * fb + (i % fb)) : 0.f;
  }
  __syncthreads();
  const int ld = fa + fb;
  for (int n = threadIdx.x; n < H; n += blockDim.x) {
    const float* w = w_init + (int64_t)n * ld + fa;
    const float b = __ldg(b_init + n);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < fb; ++j) {
      const float wj = __ldg(w + j);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(ea_s[i * fb + j], wj, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t e = e0 + i;
      if (e >= E) break;
      const float z = __ldg(P + (int64_t)__ldg(src + e) * H + n) + acc[i] + b;
      if (z0) z0[e * H + n] = z;
      h0[e * H + n] = cgr_act(z, act);
    }
  }
}

__device__ __forceinline__ float block_sum_128(float v) {
  __shared__ float ws[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0) r = (ws[0] + ws[1]) + (ws[2] + ws[3]);
  __syncthreads();
  return r;  // valid on thread 0
}

__global__ void __launch_bounds__(128) pool_ffn_kernel(const float* __restrict__ hv, const int32_t* __restrict__ atom_ptr,
                                                       const float* __restrict__ w_ffn,
                                                       const float* __restrict__ b_ffn, float* __restrict__ pooled,
                                                       float* __restrict__ out, int H) {
  const int64_t b = blockIdx.x;
  const int32_t v0 = __ldg(atom_ptr + b), v1 = __ldg(atom_ptr + b + 1);
  float part = 0.f;
  for (int n = threadIdx.x; n < H; n += blockDim.x) {
    float s = 0.f;
    for (int32_t v = v0; v < v1; ++v) s += __ldg(hv + (int64_t)v * H + n);   // ascending atom id
    pooled[b * H + n] = s;
    part = fmaf(s, __ldg(w_ffn + n), part);
  }
  const float tot = block_sum_128(part);
  if (threadIdx.x == 0) out[b] = tot + __ldg(b_ffn);
}

// dzv[v] = g[b(v)] * w_f (.) act'(zv | hv): one warp per atom, the reaction of the atom found by bisection in atom_ptr
__global__ void __launch_bounds__(256) readout_dz_kernel(const float* __restrict__ g, const int32_t* __restrict__ atom_ptr,
                                                         const float* __restrict__ w_ffn,
                                                         const float* __restrict__ hv, const float* __restrict__ zv,
                                                         int act, float* __restrict__ dzv, int64_t B, int64_t N, int H) {
  const int64_t v = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (v >= N) return;
  const int lane = threadIdx.x & 31;
  int64_t lo = 0, hi = B;                         // atom_ptr[lo] <= v < atom_ptr[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(atom_ptr + mid) <= v) lo = mid; else hi = mid;
  }
  const float gb = __ldg(g + lo);
  for (int n = lane; n < H; n += 32) {
    const int64_t idx = v * H + n;
    const float d = act == CGR_ACT_RELU ? (__ldg(hv + idx) > 0.f ? 1.f : 0.f) : cgr_act_grad(__ldg(zv + idx), 0.f, act);
    dzv[idx] = gb * __ldg(w_ffn + n) * d;
  }
}

__global__ void __launch_bounds__(128) ffn_grads_kernel(const float* __restrict__ g, const float* __restrict__ pooled,
                                                        float* __restrict__ dw, float* __restrict__ db, int64_t B,
                                                        int H) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < H) {
    float s = 0.f;
    for (int64_t b = 0; b < B; ++b) s = fmaf(__ldg(g + b), __ldg(pooled + b * H + n), s);
    dw[n] = s;
  }
  if (n == 0) {
    float s = 0.f;
    for (int64_t b = 0; b < B; ++b) s += __ldg(g + b);
    db[0] = s;
  }
}

constexpr int CS_ROWS = 32;    // rows per chunk in the column-sum pass (many small chunks: the pass is latency-bound)

__global__ void __launch_bounds__(128) colsum_pass1_kernel(const float* __restrict__ A, int64_t M, int N,
                                                           const float* __restrict__ Bm, float* __restrict__ acc,
                                                           const float* __restrict__ scale, int first,
                                                           float* __restrict__ part_col,
                                                           float* __restrict__ part_dot) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = r0 + CS_ROWS < M ? r0 + CS_ROWS : M;
  const float sc = acc ? (scale ? __ldg(scale) : 1.f) : 0.f;
  float cs = 0.f, dt = 0.f;
  if (n < N) {
    // 8 independent rows in flight per thread; sums are added in ascending row order (fixed, deterministic)
    for (int64_t r = r0; r < r1; r += 8) {
      float a[8], b[8], o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool on = r + i < r1;
        const int64_t idx = (r + i) * N + n;
        a[i] = on ? __ldg(A + idx) : 0.f;
        b[i] = (on && Bm) ? __ldg(Bm + idx) : 0.f;
        o[i] = (on && acc && !first) ? acc[idx] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        cs += a[i];
        dt = fmaf(a[i], b[i], dt);
        if (acc && r + i < r1) acc[(r + i) * N + n] = fmaf(sc, a[i], o[i]);
      }
    }
    part_col[(int64_t)blockIdx.y * N + n] = cs;
  }
  if (Bm) {
    const float tot = block_sum_128(dt);
    if (threadIdx.x == 0) part_dot[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
  }
}

// second pass: one warp per column sums the chunk partials (lanes stride over chunks, fixed shuffle tree)
__global__ void __launch_bounds__(256) colsum_pass2_kernel(const float* __restrict__ part_col,
                                                           const float* __restrict__ part_dot, int64_t chunks,
                                                           int N, int64_t n_dot, float* __restrict__ colsum,
                                                           float* __restrict__ dot) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n < N && colsum) {
    float s = 0.f;
    for (int64_t c = lane; c < chunks; c += 32) s += part_col[c * N + n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) colsum[n] = s;
  }
  if (dot && blockIdx.x == 0 && warp == 0) {
    float s = 0.f;
    for (int64_t i = lane; i < n_dot; i += 32) s += part_dot[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) dot[0] = s;
  }
}

__global__ void dropout_mask_kernel(uint64_t seed, uint32_t layer, float p, int64_t n, uint8_t* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = cgr_dropout_keep(seed, layer, (uint64_t)i, p) ? 1 : 0;
}

__global__ void mse_sum_kernel(const float* __restrict__ pred, const float* __restrict__ y, int64_t B,
                               float* __restrict__ loss, float* __restrict__ grad) {
  float s = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += 128) {
    const float d = pred[b] - y[b];
    s = fmaf(d, d, s);
    if (grad) grad[b] = 2.f * d;
  }
  const float tot = block_sum_128(s);
  if (threadIdx.x == 0 && loss) loss[0] = tot;
}

}  // namespace

int simt_splitk_choose(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = cgr_ceil_div(M, GBM) * cgr_ceil_div(N, GBN);
  if (tiles >= 148 || K < 8 * GBK) return 1;
  int64_t s = cgr_ceil_div(2 * 148, tiles);
  const int64_t max_s = K / (4 * GBK);
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  return s < 1 ? 1 : (int)s;
}

int simt_gemm(const float* A, int64_t lda, bool a_kmajor, const float* B, int64_t ldb, bool b_kmajor, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpilogue& epi, int split_k, float* partial,
              cudaStream_t st) {
  if (M <= 0 || N <= 0) return CGR_OK;
  GemmArgs g;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.epi = epi;
  if (split_k < 1) split_k = 1;
  if (split_k > 1 && !partial) {
    cgr_set_error("simt_gemm: split-K without a partial buffer");
    return CGR_ERR_ARG;
  }
  int64_t kc = cgr_ceil_div(cgr_ceil_div(K > 0 ? K : 1, split_k), GBK) * GBK;
  g.k_chunk = kc;
  split_k = (int)cgr_ceil_div(K > 0 ? K : 1, kc);
  g.partial = split_k > 1 ? partial : nullptr;
  dim3 grid((unsigned)cgr_ceil_div(N, GBN), (unsigned)cgr_ceil_div(M, GBM), (unsigned)split_k);
  CgrRange prof(epi.tag ? epi.tag : (split_k > 1 ? "sgemm_splitk" : "sgemm"), st);
  cgr_note_launch("sgemm", st, split_k > 1 ? 2 : 1);
  if (a_kmajor && b_kmajor) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else if (a_kmajor && !b_kmajor) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (!a_kmajor && b_kmajor) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
  if (split_k > 1)
    splitk_reduce_kernel<<<(unsigned)cgr_ceil_div(M * N, 256), 256, 0, st>>>(partial, split_k, M, N, C, ldc);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_gather_bonds(const float* in, const int32_t* node, const int32_t* in_ptr, const int32_t* in_idx, int flip,
                      float* out, int64_t E, int H, const GatherPost& post, cudaStream_t st) {
  if (E <= 0) return CGR_OK;
  const unsigned grid = (unsigned)cgr_ceil_div(E, 8);
  CgrRange prof("gather_bonds", st);
  cgr_note_launch("gather_bonds", st, 1);
  if (H % 4 == 0) gather_bonds_kernel<4><<<grid, 256, 0, st>>>(in, node, in_ptr, in_idx, flip, out, E, H, post);
  else gather_bonds_kernel<1><<<grid, 256, 0, st>>>(in, node, in_ptr, in_idx, flip, out, E, H, post);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_atom_sum(const float* in, const int32_t* in_ptr, const int32_t* in_idx, int flip, float* out, int64_t N,
                  int H, cudaStream_t st) {
  if (N <= 0) return CGR_OK;
  const unsigned grid = (unsigned)cgr_ceil_div(N, 8);
  CgrRange prof("atom_sum", st);
  cgr_note_launch("atom_sum", st, 1);
  if (H % 4 == 0) atom_sum_kernel<4><<<grid, 256, 0, st>>>(in, in_ptr, in_idx, flip, out, N, H);
  else atom_sum_kernel<1><<<grid, 256, 0, st>>>(in, in_ptr, in_idx, flip, out, N, H);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_edge_init(const float* P, const float* ea, const int32_t* src, const float* w_init, const float* b_init,
                   int64_t E, int fa, int fb, int H, int act, float* h0, float* z0, cudaStream_t st) {
  if (E <= 0) return CGR_OK;
  const size_t smem = (size_t)(4 * (fb > 0 ? fb : 1)) * sizeof(float);
  CgrRange prof("edge_init", st);
  cgr_note_launch("edge_init", st, 1);
  edge_init_kernel<<<(unsigned)cgr_ceil_div(E, 4), 128, smem, st>>>(P, ea, src, w_init, b_init, E, fa, fb, H, act,
                                                                    h0, z0);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_pool_ffn(const float* hv, const int32_t* atom_ptr, const float* w_ffn, const float* b_ffn, float* pooled,
                  float* out, int64_t B, int H, cudaStream_t st) {
  if (B <= 0) return CGR_OK;
  CgrRange prof("pool_ffn", st);
  cgr_note_launch("pool_ffn", st, 1);
  pool_ffn_kernel<<<(unsigned)B, 128, 0, st>>>(hv, atom_ptr, w_ffn, b_ffn, pooled, out, H);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_readout_dz(const float* g, const int32_t* atom_ptr, const float* w_ffn, const float* hv, const float* zv,
                    int act, float* dzv, int64_t B, int64_t N, int H, cudaStream_t st) {
  if (B <= 0 || N <= 0) return CGR_OK;
  CgrRange prof("readout_dz", st);
  cgr_note_launch("readout_dz", st, 1);
  readout_dz_kernel<<<(unsigned)cgr_ceil_div(N, 8), 256, 0, st>>>(g, atom_ptr, w_ffn, hv, zv, act, dzv, B, N, H);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_ffn_grads(const float* g, const float* pooled, float* dw_ffn, float* db_ffn, int64_t B, int H,
                   cudaStream_t st) {
  CgrRange prof("ffn_grads", st);
  cgr_note_launch("ffn_grads", st, 1);
  ffn_grads_kernel<<<(unsigned)cgr_ceil_div(H, 128), 128, 0, st>>>(g, pooled, dw_ffn, db_ffn, B, H);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_expand_dst(const float* ds, const int32_t* dst, float* dz, int64_t E, int H, const GatherPost& post,
                    cudaStream_t st) {
  if (E <= 0) return CGR_OK;
  CgrRange prof("expand_dst", st);
  cgr_note_launch("expand_dst", st, 1);
  expand_dst_kernel<<<(unsigned)cgr_ceil_div(E, 8), 256, 0, st>>>(ds, dst, dz, E, H, post);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

size_t simt_colsum_workspace(int64_t M, int N) {
  const int64_t chunks = cgr_ceil_div(M > 0 ? M : 1, CS_ROWS);
  return (size_t)(chunks * N + chunks * cgr_ceil_div(N, 128));
}

int simt_colsum(const float* A, int64_t M, int N, float* colsum, const float* Bm, float* dot, float* acc,
                const float* scale, bool first, float* workspace, cudaStream_t st) {
  const int64_t chunks = cgr_ceil_div(M > 0 ? M : 1, CS_ROWS);
  const int ncb = (int)cgr_ceil_div(N, 128);
  float* part_col = workspace;
  float* part_dot = workspace + chunks * N;
  dim3 grid((unsigned)ncb, (unsigned)chunks);
  CgrRange prof("colsum", st);
  cgr_note_launch("colsum", st, 2);
  colsum_pass1_kernel<<<grid, 128, 0, st>>>(A, M, N, Bm, acc, scale, first ? 1 : 0, part_col, part_dot);
  colsum_pass2_kernel<<<(unsigned)cgr_ceil_div(N, 8), 256, 0, st>>>(part_col, part_dot, chunks, N, chunks * ncb, colsum,
                                                                    Bm ? dot : nullptr);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_dropout_mask(uint64_t seed, uint32_t layer, float p, int64_t n, uint8_t* mask, cudaStream_t st) {
  if (n <= 0) return CGR_OK;
  dropout_mask_kernel<<<(unsigned)cgr_ceil_div(n, 256), 256, 0, st>>>(seed, layer, p, n, mask);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

int simt_mse_sum(const float* pred, const float* y, int64_t B, float* loss, float* grad, cudaStream_t st) {
  mse_sum_kernel<<<1, 128, 0, st>>>(pred, y, B, loss, grad);
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}
