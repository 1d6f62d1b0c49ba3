// SIMT fp32 building blocks of the layer-wise engine (CGR_ENGINE_SIMT): exact-fp32 GEMM with a
// fused epilogue, CSR gathers, deterministic reductions.  Declarations shared by api.cu.
#pragma once
#include "common.cuh"

struct GemmEpilogue {
  const float* bias = nullptr;    // [N] added to every row
  const float* res = nullptr;     // [M, ldr] residual term, scaled by *res_scale (or 1)
  int64_t ldr = 0;
  const float* res_scale = nullptr;
  float* preact = nullptr;        // [M, ldc] optional copy of the pre-activation
  int act = CGR_ACT_IDENTITY;
  float dropout_p = 0.f;          // applied iff > 0
  uint64_t seed = 0;
  uint32_t layer = 0;
  const char* tag = nullptr;      // profiling range name
};

// C[M,N] = sum_k A(m,k) * B(n,k); A(m,k) = a_kmajor ? A[m*lda+k] : A[k*lda+m], same for B.
// split_k > 1 needs `partial` (split_k*M*N floats) and ignores the epilogue except act==identity.
int simt_gemm(const float* A, int64_t lda, bool a_kmajor, const float* B, int64_t ldb, bool b_kmajor,
              float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpilogue& epi,
              int split_k, float* partial, cudaStream_t st);

// out[e] = sum_{j in in(node(e))} in[j ^ flip] - in[e ^ 1];  node = src (forward) or dst (backward)
struct GatherPost {
  int mode = 0;                   // 0: none; 1: out *= fprime; 2: out = (out + add) * fprime
  const float* h_next = nullptr;  // post-activation tensor the derivative is taken at (relu)
  const float* z = nullptr;       // pre-activation (silu / gelu)
  const float* add = nullptr;
  int act = CGR_ACT_RELU;
  float dropout_p = 0.f;
  uint64_t seed = 0;
  uint32_t layer = 0;
};
int simt_gather_bonds(const float* in, const int32_t* node, const int32_t* in_ptr, const int32_t* in_idx,
                      int flip, float* out, int64_t E, int H, const GatherPost& post, cudaStream_t st);

// s[v] = sum_{j in in(v)} in[j ^ flip]
int simt_atom_sum(const float* in, const int32_t* in_ptr, const int32_t* in_idx, int flip, float* out,
                  int64_t N, int H, cudaStream_t st);

// h0 = act(P[src e] + ea[e] . W_e^T + b);  W_e = w_init[:, fa:fa+fb] (row stride fa+fb)
int simt_edge_init(const float* P, const float* ea, const int32_t* src, const float* w_init,
                   const float* b_init, int64_t E, int fa, int fb, int H, int act, float* h0, float* z0,
                   cudaStream_t st);

// out[b] = w_f . pooled[b] + b_f,  pooled[b] = sum_{v in b} hv[v]  (ascending v)
int simt_pool_ffn(const float* hv, const int32_t* atom_ptr, const float* w_ffn, const float* b_ffn,
                  float* pooled, float* out, int64_t B, int H, cudaStream_t st);

// dzv[v] = g[b(v)] * w_f (.) act'(zv | hv)
int simt_readout_dz(const float* g, const int32_t* atom_ptr, const float* w_ffn, const float* hv,
                    const float* zv, int act, float* dzv, int64_t B, int64_t N, int H, cudaStream_t st);
// dw_f[n] = sum_b g[b] pooled[b,n];  db_f = sum_b g[b]
int simt_ffn_grads(const float* g, const float* pooled, float* dw_ffn, float* db_ffn, int64_t B, int H,
                   cudaStream_t st);

// dz[e] = ds[dst e] (.) fprime(layer)
int simt_expand_dst(const float* ds, const int32_t* dst, float* dz, int64_t E, int H, const GatherPost& post,
                    cudaStream_t st);

// colsum[n] = sum_m A[m,n];  dot = sum_{m,n} A[m,n]*Bm[m,n] (optional);
// acc[m,n] = (first ? 0 : acc[m,n]) + scale * A[m,n] (optional).  workspace: colsum_workspace floats.
size_t simt_colsum_workspace(int64_t M, int N);
int simt_colsum(const float* A, int64_t M, int N, float* colsum, const float* Bm, float* dot, float* acc,
                const float* scale, bool first, float* workspace, cudaStream_t st);

int simt_splitk_choose(int64_t M, int64_t N, int64_t K);

int simt_dropout_mask(uint64_t seed, uint32_t layer, float p, int64_t n, uint8_t* mask, cudaStream_t st);
int simt_mse_sum(const float* pred, const float* y, int64_t B, float* loss, float* grad, cudaStream_t st);
