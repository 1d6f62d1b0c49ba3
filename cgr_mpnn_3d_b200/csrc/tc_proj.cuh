// Persistent atom projection of the tcgen05 engine:  [P' | Q'] = x [W_x ; W_ox]^T + [b_i | b_o]
// (reference cgr_mpnn_3D/models/GNN.py:86 and :106-107, the parts that multiply atom features).
//
// One CTA per SM walks the work units (row tile, BN-column slice) of the launch in a fixed round-robin order --
// the slices of one row tile run at the same time on neighbouring CTAs, so the activation tile is fetched from HBM
// once and shared in L2.  The accumulator is double-buffered in TMEM (2 x 256 columns): dedicated epilogue warps
// drain unit i (TMEM -> registers -> one 128-byte output line per thread and 32-column chunk) while the MMA warp
// is already in unit i + 1, and barrier set-up, TMEM allocation and descriptor prefetch are paid once per SM instead
// of once per unit.  The one-unit-per-CTA kernel (tc_gemm_kernel<208, EPI_PLAIN>) spent about half of every CTA's
// life outside its MMAs (ncu: tensor pipe 43 % active, 656 CTAs = 4.4 waves).
//
// Same arithmetic as tc_gemm_kernel (FP16x3: A_lo.B_hi + A_hi.B_lo + A_hi.B_hi per 16-wide k-step, fp32 accumulation
// in the same order), so the results are bit-identical to it.
//
// Warp roles: 0 = TMA producer, 1 = tcgen05.mma issuer, 2 = TMEM allocation, 4..7 = epilogue (one per TMEM lane quarter).
#pragma once
#include "tc_gemm.cuh"

namespace tcp {

using tcg::A_BYTES;
using tcg::BK;
using tcg::TM;
using tcg::TcGemmParams;
using tcg::GemmBatch;

constexpr int THREADS = 256;
constexpr int SLOT_COLS = 256;                                  // TMEM columns per accumulator slot
constexpr int CH = 32;                                          // columns the epilogue reads per tcgen05.wait::ld
constexpr int AUX_BYTES = 256;
constexpr int MAX_STAGES = 4;

// What bounds this kernel is the tensor pipe fed from shared memory: an SS-mode MMA re-reads both operands from shared
// memory (9-10.5 KB per 128 x N x 16 MMA) and the FP16x3 split issues three per k-step, so a k-chunk of 12 MMAs takes
// ~2 k cycles where the math alone needs ~1 k (ncu: tensor pipe 51 % active; 1.0 PFLOP/s executed against 1.4-1.66 for
// cuBLAS bf16).  Measured equal or worse, i.e. NOT the bound: 160-wide slices with three 72 KB stages vs 208 with two
// 85 KB stages (92 us per group of 20 cfg-2 batches either way), weight multicast over clusters of 2 / 4 CTAs (104.3 /
// 104.7 / 116.6 us on a slower box), two row tiles per unit sharing every weight chunk (28 % fewer bytes from L2:
// 124 vs 105 us -- the accumulators then take both TMEM slots and the 10 k-cycle drain is exposed).  BN = 160 is
// preferred where it pads the output width less.  The lever left is cta_group::2 (each CTA reads half of B).
template <int BN_>
struct PCfg {
  static constexpr int BN = BN_;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // A (hi, lo) + B (hi, lo) of one 64-wide k-chunk
  static constexpr int FIT = (232448 - 1024 - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = FIT > MAX_STAGES ? MAX_STAGES : FIT;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + AUX_BYTES;
  static_assert(BN % 16 == 0 && BN <= SLOT_COLS && STAGES >= 2, "bad slice width");
};

struct Aux {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Aux) <= AUX_BYTES, "Aux too large");

// which batch of a group launch a global row tile belongs to (p.n_batches == 0: the launch's own operand)
struct UnitRef {
  const CUtensorMap* mapA_hi;
  const CUtensorMap* mapA_lo;
  float* out;
  int* overflow;
  int m_rows;
  int tile;        // row tile inside its batch
  bool first;      // first row tile of the batch
};
__device__ __forceinline__ UnitRef unit_ref(const TcGemmParams& p, int tile_g) {
  UnitRef u;
  if (p.n_batches > 0) {
    int bi = 0;
    while (bi + 1 < p.n_batches && tile_g >= p.gb[bi + 1].tile0) ++bi;
    const GemmBatch& g = p.gb[bi];
    u.mapA_hi = &g.tmA_hi; u.mapA_lo = &g.tmA_lo; u.out = g.out_f32; u.overflow = g.overflow; u.m_rows = g.m_rows;
    u.tile = tile_g - g.tile0; u.first = tile_g == g.tile0;
  } else {
    u.mapA_hi = &p.tmA_hi; u.mapA_lo = &p.tmA_lo; u.out = p.out_f32; u.overflow = p.overflow; u.m_rows = p.m_rows;
    u.tile = tile_g; u.first = tile_g == 0;
  }
  return u;
}

// p.mc > 1 (optional, CGR_AP_MC; measured no gain, off): clusters of mc CTAs take mc consecutive row tiles through the
// same column slice; every CTA loads 1 / mc of the weight chunk's rows and multicasts them to the whole cluster
// (tmB_*_mc: box of BN / mc rows).  n_units then counts cluster units.
template <int BN_>
__global__ void __launch_bounds__(THREADS, 1) tc_proj_kernel(const __grid_constant__ TcGemmParams p, int n_units,
                                                             int n_slices) {
  using C = PCfg<BN_>;
  constexpr int BN = C::BN, STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_BYTES = C::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = umma::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  Aux* aux = reinterpret_cast<Aux*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t mc = p.mc > 1 ? (uint32_t)p.mc : 1u;
  const uint32_t mc_rank = mc > 1 ? umma::cluster_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << mc) - 1u);
  // unit u of this CTA's cluster -> (row tile of this CTA, column slice)
  const int u0 = (int)(blockIdx.x / mc), u_step = (int)(gridDim.x / mc);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->empty[s]), mc);       // multicast: every CTA of the cluster releases the stage
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->tmem_full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->tmem_empty[s]), 4);
    }
    umma::mbar_fence_init();
    umma::tma_prefetch_desc(&p.tmB_hi);
    umma::tma_prefetch_desc(&p.tmB_lo);
  }
  if (warp == 2) {
    umma::tmem_alloc(umma::smem_u32(&aux->tmem_base), 512);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  if (mc > 1) umma::cluster_sync_all();           // the peers' barriers exist before anything is multicast to them
  const uint32_t tmem = aux->tmem_base;
  // programmatic dependent launch: the next kernel may start its own prologue as soon as SMs free up
  if (threadIdx.x == 0) umma::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      umma::grid_dep_wait();                     // x (hi, lo) may be the output of the previous kernel
      uint32_t g = 0;                            // k-chunks issued so far (ring position)
      for (int u = u0; u < n_units; u += u_step) {
        const int tg = u / n_slices, n0 = (u - tg * n_slices) * BN;
        const UnitRef r = unit_ref(p, tg * (int)mc + (int)mc_rank);
        for (int kc = 0; kc < p.num_k; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          umma::mbar_wait(umma::smem_u32(&aux->empty[s]), ph ^ 1u);
          const uint32_t full = umma::smem_u32(&aux->full[s]);
          const uint32_t st = base + s * STAGE_BYTES;
          umma::mbar_arrive_expect_tx(full, p.fast ? A_BYTES + B_BYTES : STAGE_BYTES);
          umma::tma_load_2d(r.mapA_hi, full, st, kc * BK, r.tile * TM);
          if (!p.fast) umma::tma_load_2d(r.mapA_lo, full, st + A_BYTES, kc * BK, r.tile * TM);
          if (mc > 1) {
            // this CTA's share of the weight rows, delivered to every CTA of the cluster (same stage, same offset)
            const uint32_t rows = (uint32_t)BN / mc, roff = mc_rank * rows;
            umma::tma_load_2d_mc(&p.tmB_hi_mc, full, st + 2 * A_BYTES + roff * (BK * 2), kc * BK, n0 + (int)roff, mc_mask);
            if (!p.fast)
              umma::tma_load_2d_mc(&p.tmB_lo_mc, full, st + 2 * A_BYTES + B_BYTES + roff * (BK * 2), kc * BK, n0 + (int)roff,
                                   mc_mask);
          } else {
            umma::tma_load_2d(&p.tmB_hi, full, st + 2 * A_BYTES, kc * BK, n0);
            if (!p.fast) umma::tma_load_2d(&p.tmB_lo, full, st + 2 * A_BYTES + B_BYTES, kc * BK, n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int u = u0; u < n_units; u += u_step, ++it) {
        const int n0 = (u % n_slices) * BN;
        int n_eff = p.n_total - n0;                                // columns this slice owns, rounded to the MMA granularity
        n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
        const uint32_t idesc = umma::idesc_f16_f32(TM, n_eff);
        const uint32_t slot = it & 1u;
        umma::mbar_wait(umma::smem_u32(&aux->tmem_empty[slot]), ((it >> 1) & 1u) ^ 1u);
        umma::tc_fence_after_sync();
        const uint32_t acc = tmem + slot * SLOT_COLS;
        for (int kc = 0; kc < p.num_k; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          umma::mbar_wait(umma::smem_u32(&aux->full[s]), ph);
          umma::tc_fence_after_sync();
          const uint32_t st = base + s * STAGE_BYTES;
          const uint64_t da_hi = umma::smem_desc_k_sw128(st);
          const uint64_t da_lo = umma::smem_desc_k_sw128(st + A_BYTES);
          const uint64_t db_hi = umma::smem_desc_k_sw128(st + 2 * A_BYTES);
          const uint64_t db_lo = umma::smem_desc_k_sw128(st + 2 * A_BYTES + B_BYTES);
          const int k_left = p.k_total - kc * BK;                  // K tail: skip k-steps that are all zero padding
          const int ksteps = k_left >= BK ? BK / 16 : (k_left + 15) / 16;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            if (ks >= ksteps) break;
            const uint64_t adv = (uint64_t)(ks * 32 >> 4);         // 16 fp16 = 32 bytes along K inside the swizzle row
            if (p.fast) {
              umma::mma_f16_ss(acc, da_hi + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
            } else {
              umma::mma_f16_ss(acc, da_lo + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
              umma::mma_f16_ss(acc, da_hi + adv, db_lo + adv, idesc, 1u);
              umma::mma_f16_ss(acc, da_hi + adv, db_hi + adv, idesc, 1u);
            }
          }
          if (mc > 1) umma::mma_commit_mc(umma::smem_u32(&aux->empty[s]), mc_mask);
          else umma::mma_commit(umma::smem_u32(&aux->empty[s]));   // frees the stage when these MMAs retire
          if (kc == p.num_k - 1) umma::mma_commit(umma::smem_u32(&aux->tmem_full[slot]));
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                                        // TMEM lane quarter = the 32 rows this warp owns
    const float us = __ldg(p.unscale);
    uint32_t it = 0;
    for (int u = u0; u < n_units; u += u_step, ++it) {
      const int tg = u / n_slices, n0 = (u - tg * n_slices) * BN;
      const UnitRef r = unit_ref(p, tg * (int)mc + (int)mc_rank);
      // the atom projection opens a forward: clear the per-forward overflow bit (bit 1, feature overflow, belongs to
      // the batch and stays)
      if (r.first && n0 == 0 && r.overflow && threadIdx.x == 128) atomicAnd(r.overflow, ~1);
      const uint32_t slot = it & 1u;
      umma::mbar_wait(umma::smem_u32(&aux->tmem_full[slot]), (it >> 1) & 1u);
      umma::tc_fence_after_sync();
      const uint32_t acc = tmem + slot * SLOT_COLS + ((uint32_t)(q * 32) << 16);
      const int row = r.tile * TM + q * 32 + lane;                 // this thread's row inside the batch
      const int n_cols = p.n_total - n0 < BN ? p.n_total - n0 : BN;
      float* orow = r.out + (int64_t)row * p.ldc + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < n_cols; c0 += CH) {
        float v[CH / 8][8];
#pragma unroll
        for (int g = 0; g < CH / 8; ++g)                             // all loads of the chunk in flight, one wait
          if (c0 + g * 8 < n_cols) umma::tmem_ld_x8(acc + (uint32_t)(c0 + g * 8), v[g]);
        umma::tmem_ld_wait();
        // a thread owns a row: 8 x 16 bytes = one full 128-byte line of the output per chunk
        if (row < r.m_rows) {
#pragma unroll
          for (int g = 0; g < CH / 8; ++g) {
            if (c0 + g * 8 < n_cols) {
              const int c = c0 + g * 8;
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (p.bias) { b0 = tcg::ld4(p.bias + n0 + c); b1 = tcg::ld4(p.bias + n0 + c + 4); }
              float4* dst = reinterpret_cast<float4*>(orow + c);
              dst[0] = make_float4(v[g][0] * us + b0.x, v[g][1] * us + b0.y, v[g][2] * us + b0.z, v[g][3] * us + b0.w);
              dst[1] = make_float4(v[g][4] * us + b1.x, v[g][5] * us + b1.y, v[g][6] * us + b1.z, v[g][7] * us + b1.w);
            }
          }
        }
      }
      umma::tc_fence_before_sync();                                // accumulator drained: the MMA warp may reuse it
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(umma::smem_u32(&aux->tmem_empty[slot]));
    }
  }

  __syncthreads();
  if (mc > 1) umma::cluster_sync_all();           // no CTA leaves while a peer may still multicast to it or arrive on its barriers
  if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs takes two consecutive row tiles through one column slice as
// ONE M = 256 MMA per k-step.  Each CTA stages its own 128 rows of x and only HALF of the weight slice's rows; the
// tensor cores of both SMs read the two halves, so a CTA's shared-memory operand traffic per MMA drops from
// 128 + BN to 128 + BN / 2 rows -- the SS-mode MMAs of the one-CTA kernel run at about half the math rate on it --
// and a stage shrinks to 52 KB (four fit).  The leader CTA issues the MMAs; the peer relays "my chunk has landed" to
// the leader's barrier; tcgen05.commit multicasts stage-free / accumulator-ready to both CTAs; each CTA's epilogue
// warps drain their own 128 rows and report to the leader.  Same arithmetic per output element: bit-identical.
// ---------------------------------------------------------------------------------------------------------------
template <int BN_>
struct P2Cfg {
  static constexpr int BN = BN_;
  static constexpr int HB = BN / 2;                               // weight rows each CTA of the pair stages
  static constexpr int B_BYTES = HB * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // A (hi, lo) + half of B (hi, lo)
  static constexpr int FIT = (232448 - 1024 - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = FIT > MAX_STAGES ? MAX_STAGES : FIT;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + AUX_BYTES;
  static_assert(BN % 32 == 0 && BN <= SLOT_COLS && HB % 8 == 0 && STAGE_BYTES % 1024 == 0, "bad slice width");
};
struct Aux2 {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t peer_full[MAX_STAGES];   // leader: the peer's chunk has landed in the peer's shared memory
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];           // leader: 8 arrivals (the epilogue warps of both CTAs)
  uint32_t tmem_base;
};
static_assert(sizeof(Aux2) <= AUX_BYTES, "Aux2 too large");

template <int BN_>
__global__ void __launch_bounds__(THREADS, 1) tc_proj_cg2_kernel(const __grid_constant__ TcGemmParams p, int n_units,
                                                                 int n_slices) {
  using C = P2Cfg<BN_>;
  constexpr int BN = C::BN, HB = C::HB, STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_BYTES = C::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = umma::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  Aux2* aux = reinterpret_cast<Aux2*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = umma::cluster_ctarank();
  const bool leader = rank == 0;
  const int u0 = (int)(blockIdx.x >> 1), u_step = (int)(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->empty[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->peer_full[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->tmem_full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->tmem_empty[s]), 8);
    }
    umma::mbar_fence_init();
    umma::tma_prefetch_desc(&p.tmB_hi_mc);
    umma::tma_prefetch_desc(&p.tmB_lo_mc);
  }
  if (warp == 2) {
    umma::tmem_alloc_cg2(umma::smem_u32(&aux->tmem_base), 512);
    umma::tmem_relinquish_cg2();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  umma::cluster_sync_all();                       // both CTAs' barriers and TMEM exist before either is used
  const uint32_t tmem = aux->tmem_base;
  if (threadIdx.x == 0) umma::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      umma::grid_dep_wait();                     // x (hi, lo) may be the output of the previous kernel
      uint32_t g = 0;
      for (int u = u0; u < n_units; u += u_step) {
        const int tg = u / n_slices, n0 = (u - tg * n_slices) * BN;
        const UnitRef r = unit_ref(p, 2 * tg + (int)rank);
        for (int kc = 0; kc < p.num_k; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          umma::mbar_wait_cluster(umma::smem_u32(&aux->empty[s]), ph ^ 1u);
          const uint32_t full = umma::smem_u32(&aux->full[s]);
          const uint32_t st = base + s * STAGE_BYTES;
          umma::mbar_arrive_expect_tx(full, p.fast ? A_BYTES + B_BYTES : STAGE_BYTES);
          umma::tma_load_2d(r.mapA_hi, full, st, kc * BK, r.tile * TM);
          if (!p.fast) umma::tma_load_2d(r.mapA_lo, full, st + A_BYTES, kc * BK, r.tile * TM);
          // this CTA's half of the weight slice's rows (tmB_*_mc: box of BN / 2 rows)
          umma::tma_load_2d(&p.tmB_hi_mc, full, st + 2 * A_BYTES, kc * BK, n0 + (int)rank * HB);
          if (!p.fast) umma::tma_load_2d(&p.tmB_lo_mc, full, st + 2 * A_BYTES + B_BYTES, kc * BK, n0 + (int)rank * HB);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && !leader) {
      // ---------------------------------------------------------------- peer CTA: relay "chunk landed" to the leader
      uint32_t g = 0;
      for (int u = u0; u < n_units; u += u_step)
        for (int kc = 0; kc < p.num_k; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          umma::mbar_wait(umma::smem_u32(&aux->full[s]), ph);
          umma::mbar_arrive_remote(umma::smem_u32(&aux->peer_full[s]), 0u);
        }
    } else if (lane == 0) {
      // ---------------------------------------------------------------- leader CTA: MMA issuer of the pair
      uint32_t g = 0, it = 0;
      for (int u = u0; u < n_units; u += u_step, ++it) {
        const int n0 = (u % n_slices) * BN;
        int n_eff = p.n_total - n0;                                // columns this slice owns (multiple of 32: both halves of 16)
        n_eff = n_eff >= BN ? BN : ((n_eff + 31) & ~31);
        const uint32_t idesc = umma::idesc_f16_f32(2 * TM, n_eff);
        const uint32_t slot = it & 1u;
        umma::mbar_wait_cluster(umma::smem_u32(&aux->tmem_empty[slot]), ((it >> 1) & 1u) ^ 1u);
        umma::tc_fence_after_sync();
        const uint32_t acc = tmem + slot * SLOT_COLS;
        for (int kc = 0; kc < p.num_k; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          umma::mbar_wait(umma::smem_u32(&aux->full[s]), ph);
          umma::mbar_wait_cluster(umma::smem_u32(&aux->peer_full[s]), ph);
          umma::tc_fence_after_sync();
          const uint32_t st = base + s * STAGE_BYTES;
          const uint64_t da_hi = umma::smem_desc_k_sw128(st);
          const uint64_t da_lo = umma::smem_desc_k_sw128(st + A_BYTES);
          const uint64_t db_hi = umma::smem_desc_k_sw128(st + 2 * A_BYTES);
          const uint64_t db_lo = umma::smem_desc_k_sw128(st + 2 * A_BYTES + B_BYTES);
          const int k_left = p.k_total - kc * BK;                  // K tail: skip k-steps that are all zero padding
          const int ksteps = k_left >= BK ? BK / 16 : (k_left + 15) / 16;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            if (ks >= ksteps) break;
            const uint64_t adv = (uint64_t)(ks * 32 >> 4);         // 16 fp16 = 32 bytes along K inside the swizzle row
            if (p.fast) {
              umma::mma_f16_ss_cg2(acc, da_hi + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
            } else {
              umma::mma_f16_ss_cg2(acc, da_lo + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
              umma::mma_f16_ss_cg2(acc, da_hi + adv, db_lo + adv, idesc, 1u);
              umma::mma_f16_ss_cg2(acc, da_hi + adv, db_hi + adv, idesc, 1u);
            }
          }
          umma::mma_commit_cg2_mc(umma::smem_u32(&aux->empty[s]), (uint16_t)3);   // the stage is free in both CTAs
          if (kc == p.num_k - 1) umma::mma_commit_cg2_mc(umma::smem_u32(&aux->tmem_full[slot]), (uint16_t)3);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int q = warp & 3;
    const float us = __ldg(p.unscale);
    uint32_t it = 0;
    for (int u = u0; u < n_units; u += u_step, ++it) {
      const int tg = u / n_slices, n0 = (u - tg * n_slices) * BN;
      const UnitRef r = unit_ref(p, 2 * tg + (int)rank);
      if (r.first && n0 == 0 && r.overflow && threadIdx.x == 128) atomicAnd(r.overflow, ~1);
      const uint32_t slot = it & 1u;
      umma::mbar_wait_cluster(umma::smem_u32(&aux->tmem_full[slot]), (it >> 1) & 1u);
      umma::tc_fence_after_sync();
      const uint32_t acc = tmem + slot * SLOT_COLS + ((uint32_t)(q * 32) << 16);
      const int row = r.tile * TM + q * 32 + lane;
      const int n_cols = p.n_total - n0 < BN ? p.n_total - n0 : BN;
      float* orow = r.out + (int64_t)row * p.ldc + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < n_cols; c0 += CH) {
        float v[CH / 8][8];
#pragma unroll
        for (int g = 0; g < CH / 8; ++g)
          if (c0 + g * 8 < n_cols) umma::tmem_ld_x8(acc + (uint32_t)(c0 + g * 8), v[g]);
        umma::tmem_ld_wait();
        if (row < r.m_rows) {
#pragma unroll
          for (int g = 0; g < CH / 8; ++g) {
            if (c0 + g * 8 < n_cols) {
              const int c = c0 + g * 8;
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (p.bias) { b0 = tcg::ld4(p.bias + n0 + c); b1 = tcg::ld4(p.bias + n0 + c + 4); }
              float4* dst = reinterpret_cast<float4*>(orow + c);
              dst[0] = make_float4(v[g][0] * us + b0.x, v[g][1] * us + b0.y, v[g][2] * us + b0.z, v[g][3] * us + b0.w);
              dst[1] = make_float4(v[g][4] * us + b1.x, v[g][5] * us + b1.y, v[g][6] * us + b1.z, v[g][7] * us + b1.w);
            }
          }
        }
      }
      umma::tc_fence_before_sync();                                // accumulator drained: the leader may reuse the slot
      __syncwarp();
      if (lane == 0) umma::mbar_arrive_remote(umma::smem_u32(&aux->tmem_empty[slot]), 0u);
    }
  }

  __syncthreads();
  umma::cluster_sync_all();                       // no CTA leaves while the pair's MMAs / arrivals may still touch it
  if (warp == 2) umma::tmem_dealloc_cg2(tmem, 512);
}

}  // namespace tcp
