// General FP16x3 tcgen05 GEMM for the training path:  C[M,N] = sum_k A(m,k) B(n,k)  with fp32 accumulation,
// each operand either K-major (rows of K, as activations / Linear weights are stored) or MN-major (rows of the
// contracted index, e.g. dz^T m for weight gradients) -- the layout is expressed purely through the TMA boxes
// and the UMMA shared-memory descriptors, no transposed copies are made.  Epilogue = the SIMT GEMM's
// (bias, scaled residual, pre-activation copy, activation, Philox dropout) or split-K partial sums.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace tcg2 {

constexpr int TM = 128, TN = 128, BK = 64;      // CTA tile and k-chunk (fp16 elements)
constexpr int OP_BYTES = TM * BK * 2;           // one operand half (hi or lo) of a stage: 16 KB
constexpr int STAGE_BYTES = 4 * OP_BYTES;       // A_hi, A_lo, B_hi, B_lo
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr int YP = TN + 4;                      // padded fp32 staging row
constexpr int AUX_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;
static_assert(TM * YP * 4 <= STAGES * STAGE_BYTES, "staging tile must fit in the drained pipeline");

struct Params {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  int64_t M, N, K;
  int64_t k_chunks_per_split;   // k-chunks of BK handled by one blockIdx.z
  const float* unscale_a;       // device scalars 1/scale_A, 1/scale_B (null = 1)
  const float* unscale_b;
  float* C;                     // [M, ldc]            (split-K: partial sums [z][M][N])
  int64_t ldc;
  int split_k;
  // epilogue (ignored for split-K partials)
  const float* bias;            // [N]
  const float* res;             // [M, ldr]
  int64_t ldr;
  const float* res_scale;       // device scalar or null (=1)
  float* preact;                // [M, ldc] or null
  int act;
  float dropout_p;
  uint64_t seed;
  uint32_t layer;
};

struct Aux {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

// MN-major operand, 128-byte swizzle: a stage holds 64-element MN blocks of [BK k-rows x 128 bytes]; inside a
// block 8 k-rows form a 1024-byte swizzle atom (SBO = 1024), MN blocks are LBO = BK*128 bytes apart.
// (cute make_umma_desc<Major::MN>, LayoutType::B128: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.)
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((BK * 128) >> 4) << 16;       // leading byte offset: next 64-wide MN block
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: next group of 8 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Several GEMMs of one shape in one launch (the weight gradients of all layers): blockIdx.z = problem * split_k + split
constexpr int MAX_BATCH = 8;
struct BatchParams {
  Params prob[MAX_BATCH];
  int n_prob;
};

template <bool A_MN, bool B_MN>
__device__ __forceinline__ void gemm2_body(const Params& p, const int zsplit) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = umma::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  Aux* aux = reinterpret_cast<Aux*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
  const int64_t kc_total = (p.K + BK - 1) / BK;
  const int64_t kc_beg = (int64_t)zsplit * p.k_chunks_per_split;
  int64_t kc_end = kc_beg + p.k_chunks_per_split;
  if (kc_end > kc_total) kc_end = kc_total;
  const int num_k = (int)(kc_end - kc_beg);      // >= 1 by construction of the grid

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
    umma::tmem_alloc(umma::smem_u32(&aux->tmem_base), 128);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = aux->tmem_base;

  if (warp == 0) {
    for (int i = 0; i < num_k; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
      if (lane == 0) {
        umma::mbar_wait(umma::smem_u32(&aux->empty[s]), ph ^ 1u);
        const uint32_t full = umma::smem_u32(&aux->full[s]);
        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
        const int k0 = (int)((kc_beg + i) * BK);
        umma::mbar_arrive_expect_tx(full, STAGE_BYTES);
        if (A_MN) {                                  // two 64-wide MN blocks of [BK k-rows x 64 cols]
          umma::tma_load_2d(&p.tmA_hi, full, st, (int)m0, k0);
          umma::tma_load_2d(&p.tmA_hi, full, st + OP_BYTES / 2, (int)m0 + 64, k0);
          umma::tma_load_2d(&p.tmA_lo, full, st + OP_BYTES, (int)m0, k0);
          umma::tma_load_2d(&p.tmA_lo, full, st + OP_BYTES + OP_BYTES / 2, (int)m0 + 64, k0);
        } else {                                     // [128 rows x 64 k-cols]
          umma::tma_load_2d(&p.tmA_hi, full, st, k0, (int)m0);
          umma::tma_load_2d(&p.tmA_lo, full, st + OP_BYTES, k0, (int)m0);
        }
        if (B_MN) {
          umma::tma_load_2d(&p.tmB_hi, full, st + 2 * OP_BYTES, (int)n0, k0);
          umma::tma_load_2d(&p.tmB_hi, full, st + 2 * OP_BYTES + OP_BYTES / 2, (int)n0 + 64, k0);
          umma::tma_load_2d(&p.tmB_lo, full, st + 3 * OP_BYTES, (int)n0, k0);
          umma::tma_load_2d(&p.tmB_lo, full, st + 3 * OP_BYTES + OP_BYTES / 2, (int)n0 + 64, k0);
        } else {
          umma::tma_load_2d(&p.tmB_hi, full, st + 2 * OP_BYTES, k0, (int)n0);
          umma::tma_load_2d(&p.tmB_lo, full, st + 3 * OP_BYTES, k0, (int)n0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t id = idesc(TM, TN, A_MN, B_MN);
    for (int i = 0; i < num_k; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
      if (lane == 0) {
        umma::mbar_wait(umma::smem_u32(&aux->full[s]), ph);
        umma::tc_fence_after_sync();
        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
        const uint64_t da_hi = A_MN ? smem_desc_mn_sw128(st) : umma::smem_desc_k_sw128(st);
        const uint64_t da_lo = A_MN ? smem_desc_mn_sw128(st + OP_BYTES) : umma::smem_desc_k_sw128(st + OP_BYTES);
        const uint64_t db_hi = B_MN ? smem_desc_mn_sw128(st + 2 * OP_BYTES) : umma::smem_desc_k_sw128(st + 2 * OP_BYTES);
        const uint64_t db_lo = B_MN ? smem_desc_mn_sw128(st + 3 * OP_BYTES) : umma::smem_desc_k_sw128(st + 3 * OP_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          // 16 k per MMA: K-major = 32 bytes inside the swizzle row; MN-major = 16 k-rows = two 1024-byte atoms
          const uint64_t adv_a = (uint64_t)((A_MN ? ks * 2048 : ks * 32) >> 4);
          const uint64_t adv_b = (uint64_t)((B_MN ? ks * 2048 : ks * 32) >> 4);
          umma::mma_f16_ss(tmem, da_lo + adv_a, db_hi + adv_b, id, (i | ks) ? 1u : 0u);
          umma::mma_f16_ss(tmem, da_hi + adv_a, db_lo + adv_b, id, 1u);
          umma::mma_f16_ss(tmem, da_hi + adv_a, db_hi + adv_b, id, 1u);
        }
        umma::mma_commit(umma::smem_u32(&aux->empty[s]));
        if (i == num_k - 1) umma::mma_commit(umma::smem_u32(&aux->tmem_full));
      }
      __syncwarp();
    }
  }

  umma::mbar_wait(umma::smem_u32(&aux->tmem_full), 0);
  umma::tc_fence_after_sync();
  float* y_s = reinterpret_cast<float*>(smem);                   // [TM][YP]
  {
    const float us = (p.unscale_a ? __ldg(p.unscale_a) : 1.f) * (p.unscale_b ? __ldg(p.unscale_b) : 1.f);
    const int q = warp & 3, grp = warp >> 2;
    const int row = q * 32 + lane;
#pragma unroll
    for (int cc = grp * 8; cc < TN; cc += 8 * (NWARPS / 4)) {
      float v[8];
      umma::tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      umma::tmem_ld_wait();
      float4* dst = reinterpret_cast<float4*>(y_s + row * YP + cc);
      dst[0] = make_float4(v[0] * us, v[1] * us, v[2] * us, v[3] * us);
      dst[1] = make_float4(v[4] * us, v[5] * us, v[6] * us, v[7] * us);
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();

  // coalesced row-wise epilogue: lane owns 4 consecutive columns
  const int c = 4 * lane;
  const int64_t n = n0 + c;
  if (n < p.N) {
    const bool partial = p.split_k > 1;
    float* Cb = partial ? p.C + (int64_t)zsplit * p.M * p.N : p.C;
    const int64_t ldc = partial ? p.N : p.ldc;
    const float rs = (!partial && p.res) ? (p.res_scale ? __ldg(p.res_scale) : 1.f) : 0.f;
    const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (!partial && p.bias)
      for (int i = 0; i < 4; ++i) b4[i] = n + i < p.N ? __ldg(p.bias + n + i) : 0.f;
    // rows are handled four at a time: the residual loads of all four are in flight before the first is used
    const bool vec = (n + 3 < p.N) && ((ldc & 3) == 0) && (((uintptr_t)Cb & 15) == 0) &&
                     (!p.res || ((p.ldr & 3) == 0 && ((uintptr_t)p.res & 15) == 0)) &&
                     (!p.preact || ((uintptr_t)p.preact & 15) == 0);
    for (int r0 = warp * 4; r0 < TM; r0 += NWARPS * 4) {
      float res4[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t m = m0 + r0 + u;
#pragma unroll
        for (int i = 0; i < 4; ++i) res4[u][i] = 0.f;
        if (!partial && p.res && m < p.M) {
          if (vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.res + m * p.ldr + n));
            res4[u][0] = t.x; res4[u][1] = t.y; res4[u][2] = t.z; res4[u][3] = t.w;
          } else {
            for (int i = 0; i < 4; ++i) if (n + i < p.N) res4[u][i] = __ldg(p.res + m * p.ldr + n + i);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u;
        const int64_t m = m0 + r;
        if (m >= p.M) break;
        const float4 y4 = *reinterpret_cast<const float4*>(y_s + r * YP + c);
        float y[4] = {y4.x, y4.y, y4.z, y4.w};
        float zpre[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = y[i];
          if (!partial) {
            v += b4[i];
            if (p.res) v = fmaf(rs, res4[u][i], v);
            zpre[i] = v;
            v = cgr_act(v, p.act);
            if (p.dropout_p > 0.f)
              v = cgr_dropout_keep(p.seed, p.layer, (uint64_t)(m * p.N + n + i), p.dropout_p) ? v * keep_scale : 0.f;
          }
          y[i] = v;
        }
        if (vec) {
          if (!partial && p.preact) *reinterpret_cast<float4*>(p.preact + m * ldc + n) = make_float4(zpre[0], zpre[1], zpre[2], zpre[3]);
          *reinterpret_cast<float4*>(Cb + m * ldc + n) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
          for (int i = 0; i < 4; ++i) {
            if (n + i >= p.N) break;
            if (!partial && p.preact) p.preact[m * ldc + n + i] = zpre[i];
            Cb[m * ldc + n + i] = y[i];
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 128);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm2_kernel(const __grid_constant__ Params p) {
  gemm2_body<A_MN, B_MN>(p, (int)blockIdx.z);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm2_batched_kernel(const __grid_constant__ BatchParams bp) {
  const int split_k = bp.prob[0].split_k;
  gemm2_body<A_MN, B_MN>(bp.prob[blockIdx.z / split_k], (int)(blockIdx.z % split_k));
}

}  // namespace tcg2
