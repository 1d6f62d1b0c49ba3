// The fused tcgen05 kernel of the tensor-core engine: TMA-fed FP16x3 GEMM of one 128-row tile with the
// accumulator in TMEM, followed by a shared-memory epilogue (plain / directed-bond gather / readout).
//
// Measured on B200 (tools/tc_phase_timing.py): an SS-mode tcgen05.mma with M=128, N=80 costs ~97 cycles
// (operand fetch from shared memory, not the 40-cycle math), so wide slices amortise the A-operand fetch:
// the kernel is templated on the slice width BN (80 for small batches that need many CTAs, 208 for large
// batches) and the epilogue walks the slice in column chunks of <= 128.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace tcg {

constexpr int TM = 128;                 // rows of a tile (UMMA M)
constexpr int BK = 64;                  // fp16 elements per k-chunk = one 128-byte swizzle row
constexpr int A_BYTES = TM * BK * 2;    // 16 KB
constexpr int AUX_BYTES = 4096;
constexpr int SMEM_LIMIT = 232448;      // 227 KB opt-in maximum per CTA
constexpr int RU = 4;                   // rows processed together by a warp in the epilogue (ILP)
constexpr int NBR = 8;                  // neighbour slots per row kept in the packed descriptor

// EPI_BOND_BWD / EPI_INIT_BWD: the tile-local backward of a bond layer / of the edge initialisation (training)
enum { EPI_PLAIN = 0, EPI_BOND = 1, EPI_READOUT = 2, EPI_BOND_BWD = 3, EPI_INIT_BWD = 4 };

// NT = threads per CTA.  NT=512: one CTA per SM, deep pipeline, dedicated buffer for the prefetched fp32 operand.
// NT=256: two CTAs per SM (<= 113 KB each): 2-stage pipeline, the fp32 operand is loaded into the drained pipeline
// after the GEMM -- its latency and the whole epilogue of one CTA hide behind the other CTA's MMA stream.
template <int BN_, int EPI, int NT_>
struct Cfg {
  static constexpr int NT = NT_;
  static constexpr int NWARPS = NT / 32;
  static constexpr int CTAS_PER_SM = NT <= 384 ? 2 : 1;
  static constexpr bool R_ALIAS = CTAS_PER_SM > 1;      // fp32 operand chunk 0 lives in the drained pipeline
  static constexpr int BN = BN_;                          // output columns per CTA (UMMA N), multiple of 16
  static constexpr int NCH = BN > 128 ? 2 : 1;            // epilogue column chunks
  static constexpr int CH = BN / NCH;                     // columns per chunk (80 / 104 / 128)
  static constexpr int CHP = CH + 4;                      // padded fp32 staging row (conflict-free float4 rows)
  static constexpr int VL = CH / 4;                       // lanes owning a float4 column group
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int R_BYTES = EPI == EPI_PLAIN ? 0 : TM * CH * 4;   // TMA-loaded fp32 epilogue operand (one chunk)
  static constexpr int R_DEDICATED = R_ALIAS ? 0 : R_BYTES;
  static constexpr int Y_BYTES = ((TM * CHP * 4 + 127) / 128) * 128;
  static constexpr int BUDGET = SMEM_LIMIT / CTAS_PER_SM - (CTAS_PER_SM > 1 ? 1024 : 0);
  static constexpr int FIT = (BUDGET - R_DEDICATED - AUX_BYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 4 ? 4 : FIT;
  // [B_hi ; B_lo] are adjacent in a stage, so A_hi x [B_hi;B_lo] is ONE MMA of N = 2*BN when that fits the
  // 256-column instruction limit: 2 MMAs per k-step instead of 3, and A_hi / B_hi are fetched once less
  static constexpr bool CAT = 2 * BN <= 256;
  static constexpr int ACC_COLS = CAT ? 2 * BN : BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + R_DEDICATED + AUX_BYTES + 1024;
  // backward epilogues: per-warp column sums [NWARPS][CH] behind the staging buffers
  static constexpr int RED_OFF = Y_BYTES + ((NCH > 1 || R_ALIAS) ? R_BYTES : 0);
  static_assert(EPI < 3 || RED_OFF + NWARPS * CH * 4 <= STAGES * STAGE_BYTES, "reduction scratch must fit in the pipeline buffers");
  static_assert(BN % 16 == 0 && BN <= 256 && CH % 4 == 0 && VL <= 32, "bad slice width");
  static_assert(STAGES >= 2, "pipeline needs two stages");
  static_assert(Y_BYTES + ((NCH > 1 || R_ALIAS) ? R_BYTES : 0) <= STAGES * STAGE_BYTES, "epilogue staging must fit in the pipeline buffers");
  static_assert(!(R_ALIAS && NCH > 1), "aliased fp32 operand supports a single column chunk");
  static_assert(SMEM_BYTES <= BUDGET, "shared memory budget");
  static_assert(TMEM_COLS * CTAS_PER_SM <= 512, "TMEM budget");
};

constexpr int MAX_GEMM_GROUP = 24;        // batches one EPI_PLAIN launch can take (atom projection of a group forward)
// EPI_PLAIN over several batches in one launch: row tiles of batch i are blockIdx.y in [tile0, next tile0)
struct GemmBatch {
  CUtensorMap tmA_hi, tmA_lo;
  float* out_f32;
  int* overflow;
  int m_rows;
  int tile0;
};

struct TcGemmParams {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  CUtensorMap tmR;              // fp32 [rows, cols] operand of the epilogue (BOND: h0, READOUT: Q), box = [128, CH]
  int num_k;                    // k-chunks of BK
  int k_total;                  // real K: the last chunk only issues the k-steps that hold data
  int n_total;                  // real output columns
  int m_rows;                   // real rows (EPI_PLAIN)
  int r_col0;                   // column offset of the R operand inside its tensor (READOUT: H)
  const float* unscale;         // device scalar 1 / (scale_A * scale_W)
  const float* bias;            // [n_total] or null
  // EPI_PLAIN
  float* out_f32;
  int64_t ldc;
  int n_batches;                // > 0: the A operand, rows and output of every batch come from gb[] (group launch)
  int mc;                       // EPI_PLAIN: > 1 = clusters of mc row tiles (grid.y) share the weight chunks -- every CTA
                                // loads 1/mc of the B rows and multicasts them (tmB_*_mc: box of BN / mc rows)
  CUtensorMap tmB_hi_mc, tmB_lo_mc;
  GemmBatch gb[MAX_GEMM_GROUP];
  // tile-local epilogues
  const int32_t* tile_info;     // [T][8]: ebase, ecount, abase, acount, rx0, rxcount, 0, 0
  const int32_t* in_ptr;
  const int32_t* in_idx;
  const int32_t* src;
  const int32_t* atom_ptr;
  const float* skip;            // device scalar or null (=1)
  int act;
  float dropout_p;
  uint64_t seed;
  uint32_t layer;
  __half* o_hi;                 // next operand, [T*128, ldo]
  __half* o_lo;
  int64_t ldo;
  const float* w_ffn;
  const float* b_ffn;
  float* partial_out;           // [n_slices, B]
  float* out;                   // [B] final energies, written by the last slice CTA of each tile
  int* tile_counter;            // [T] arrival counters of the readout (self-resetting)
  int64_t n_rxn;
  int* overflow;                // sticky flag: an activation left the fp16 range
  float* hv_out;                // READOUT, training: hv [N, H] fp32 (ReLU mask of the readout backward)
  // ---- backward epilogues (EPI_BOND_BWD / EPI_INIT_BWD) ----
  const __half* mask_hi;        // hi part of the layer's saved output (tile-packed): dz = dh * [h > 0] * keep_scale
  int64_t ld_mask;
  float keep_scale;             // 1 / (1 - dropout_p) of the layer
  float* dh0_acc;               // BOND_BWD: [T*128, H] fp32, += skip * dz  (INIT_BWD reads it through tmR)
  int dh0_first;                // first contribution: store instead of accumulate
  float* colsum_partial;        // [T][H]   per-tile column sums of dz  (bias gradient)
  float* dskip_partial;         // [T][n_slices * NCH]  per-CTA sum of dz . h0 (skip-weight gradient) or null
  const unsigned int* gamax_in;   // amax (float bits) that fixes the power-of-two scale of the A operand
  const unsigned int* gamax_out;  // amax that fixes the scale of the produced operand
  unsigned int* gamax_track;      // true amax of the produced operand (fixes the next kernel's scale)
  float* gunscale_out;            // 1 / scale of the produced operand, for the weight-gradient GEMMs
  float* dz0_out;               // INIT_BWD: dz_0 [E, H] fp32 in bond order (bond-feature weight gradient)
  long long* dbg;               // optional [n_cta][8] clock64 stamps of the kernel phases (debug)
  int fast;                     // 1: single-pass fp16 (only the hi halves are loaded and multiplied): "fast" precision mode
};

struct Aux {                    // small per-CTA shared state, lives after the pipeline buffers
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t tmem_full;
  uint64_t r_full[2];
  uint32_t tmem_base;
  int32_t info[8];
  uint16_t ptr_l[TM + 2];       // local CSR offsets of the tile's atoms
  uint8_t src_l[TM];            // local source atom of each bond row
  uint8_t idx_l[TM];            // local bond ids grouped by target atom
  float tat[TM];                // readout: per-atom dot with w_ffn
  // per epilogue row (BOND: bond j -> in-bonds of src(j); READOUT: atom v -> in-bonds of v):
  uint2 nbr[TM];                // first NBR neighbour rows, one byte each
  uint16_t nbr_pb[TM];          // CSR offset of the row's neighbour list (for degrees > NBR)
  uint8_t nbr_deg[TM];          // its length
  float red2[16];               // backward: per-warp partial of the skip-weight gradient
};
static_assert(sizeof(Aux) <= AUX_BYTES, "Aux too large");

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void add4(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
template <bool RELU>
__device__ __forceinline__ float act_t(float z, int act) { return RELU ? fmaxf(z, 0.f) : cgr_act(z, act); }

// Power-of-two scale that maps a gradient tensor of magnitude `amax` into [2^7, 2^8] before its FP16 (hi, lo) split:
// far from the fp16 overflow (2^16) when the next tensor grows, and ~2^30 above the lo part's underflow.
__device__ __forceinline__ float grad_scale(unsigned int amax_bits) {
  const float a = __uint_as_float(amax_bits);
  if (!(a > 0.f) || !isfinite(a)) return 1.f;
  int e;
  frexpf(a, &e);
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return ldexpf(1.f, 8 - e);
}

// sum of the staged rows named by the packed neighbour descriptors of RU consecutive epilogue rows
template <int CHP>
__device__ __forceinline__ void nbr_sum(const float* y_s, const Aux* aux, int r0, int nrows, int c, float4 (&acc)[RU]) {
  uint2 nb[RU];
  int deg[RU], maxdeg = 0;
#pragma unroll
  for (int u = 0; u < RU; ++u) {
    const int r = r0 + u < nrows ? r0 + u : r0;
    nb[u] = aux->nbr[r];
    deg[u] = r0 + u < nrows ? (int)aux->nbr_deg[r] : 0;
    maxdeg = deg[u] > maxdeg ? deg[u] : maxdeg;
    acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (t < maxdeg) {
#pragma unroll
      for (int u = 0; u < RU; ++u)
        if (t < deg[u]) add4(acc[u], ld4(y_s + (int)((nb[u].x >> (8 * t)) & 0xffu) * CHP + c));
    }
  }
  const int fast = maxdeg < NBR ? maxdeg : NBR;
  for (int t = 4; t < fast; ++t) {
#pragma unroll
    for (int u = 0; u < RU; ++u)
      if (t < deg[u]) add4(acc[u], ld4(y_s + (int)((nb[u].y >> (8 * (t - 4))) & 0xffu) * CHP + c));
  }
  if (maxdeg > NBR) {
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      if (r0 + u >= nrows) break;
      const int pb = aux->nbr_pb[r0 + u];
      for (int t = NBR; t < deg[u]; ++t) add4(acc[u], ld4(y_s + (int)aux->idx_l[pb + t] * CHP + c));
    }
  }
}

// scaled FP16 (hi, lo) split of one float4 column group, stored as two 8-byte words
__device__ __forceinline__ void store_split4(const float4 v, float scale, __half* hi, __half* lo) {
  const float z0 = v.x * scale, z1 = v.y * scale, z2 = v.z * scale, z3 = v.w * scale;
  const __half2 hi01 = __floats2half2_rn(z0, z1), hi23 = __floats2half2_rn(z2, z3);
  const float2 f01 = __half22float2(hi01), f23 = __half22float2(hi23);
  const __half2 lo01 = __floats2half2_rn(z0 - f01.x, z1 - f01.y), lo23 = __floats2half2_rn(z2 - f23.x, z3 - f23.y);
  uint2 ph, pl;
  ph.x = *reinterpret_cast<const uint32_t*>(&hi01); ph.y = *reinterpret_cast<const uint32_t*>(&hi23);
  pl.x = *reinterpret_cast<const uint32_t*>(&lo01); pl.y = *reinterpret_cast<const uint32_t*>(&lo23);
  *reinterpret_cast<uint2*>(hi) = ph;
  *reinterpret_cast<uint2*>(lo) = pl;
}
__device__ __forceinline__ float amax4(const float4 v) {
  return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
}

// RELU: compile-time fast path for the reference's default activation (branch-free epilogue)
template <int BN_, int EPI, bool RELU, int NT_>
__global__ void __launch_bounds__(NT_, (NT_ <= 384 ? 2 : 1)) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
  using C = Cfg<BN_, EPI, NT_>;
  constexpr int THREADS = C::NT, NWARPS = C::NWARPS;
  constexpr int BN = C::BN, CH = C::CH, CHP = C::CHP, VL = C::VL, NCH = C::NCH, STAGES = C::STAGES;
  constexpr int STAGE_BYTES = C::STAGE_BYTES, B_BYTES = C::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = umma::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  // fp32 operand chunk 0: dedicated buffer (prefetched during the GEMM) or, with two CTAs per SM, the drained pipeline
  float* r0_s = reinterpret_cast<float*>(C::R_ALIAS ? smem + C::Y_BYTES : smem + STAGES * STAGE_BYTES);
  Aux* aux = reinterpret_cast<Aux*>(smem + STAGES * STAGE_BYTES + C::R_DEDICATED);
  float* y_s = reinterpret_cast<float*>(smem);                                  // [TM][CHP], aliases the drained pipeline
  float* r1_s = reinterpret_cast<float*>(smem + C::Y_BYTES);                    // [TM][CH], chunk 1 (alias)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x;                                  // slices of one tile are neighbours: they share A in L2
  int tile = blockIdx.y;
  // operand / output of this CTA's rows: the launch's own, or (group launch of the atom projection) its batch's
  const CUtensorMap* mapA_hi = &p.tmA_hi;
  const CUtensorMap* mapA_lo = &p.tmA_lo;
  float* out_f32 = p.out_f32;
  int* overflow = p.overflow;
  int m_rows = p.m_rows;
  bool first_cta = blockIdx.x == 0 && blockIdx.y == 0;
  if (EPI == EPI_PLAIN && p.n_batches > 0) {
    int bi = 0;
    while (bi + 1 < p.n_batches && tile >= p.gb[bi + 1].tile0) ++bi;
    const GemmBatch& gbi = p.gb[bi];
    first_cta = blockIdx.x == 0 && tile == gbi.tile0;
    tile -= gbi.tile0;
    mapA_hi = &gbi.tmA_hi; mapA_lo = &gbi.tmA_lo;
    out_f32 = gbi.out_f32; overflow = gbi.overflow; m_rows = gbi.m_rows;
  }
  const uint32_t mc = (EPI == EPI_PLAIN && p.mc > 1) ? (uint32_t)p.mc : 1u;
  const uint32_t mc_rank = mc > 1 ? umma::cluster_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << mc) - 1u);
  const int n0 = slice * BN;
  // columns this slice really owns, rounded up to the MMA granularity (runtime N of the instruction)
  int n_eff = p.n_total - n0;
  n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
  long long* dbg = p.dbg ? p.dbg + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
#ifdef CGR_FWD_STAMPS      // debug builds only: a predicated-off stamp still waits on the scoreboards it shares (see tc_fwd.cuh)
#define TC_STAMP(k) do { if (dbg && threadIdx.x == 64) dbg[k] = clock64(); } while (0)
#else
#define TC_STAMP(k) do { (void)dbg; } while (0)
#endif
  TC_STAMP(0);
  // the atom projection opens a forward: clear the per-forward overflow bit (bit 1, feature overflow, belongs to the batch)
  if (EPI == EPI_PLAIN && overflow && first_cta && threadIdx.x == 0) atomicAnd(overflow, ~1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->empty[s]), mc);       // multicast: every CTA of the cluster releases the stage
    }
    umma::mbar_init(umma::smem_u32(&aux->tmem_full), 1);
    umma::mbar_init(umma::smem_u32(&aux->r_full[0]), 1);
    umma::mbar_init(umma::smem_u32(&aux->r_full[1]), 1);
    umma::mbar_fence_init();
    umma::tma_prefetch_desc(mapA_hi);
    umma::tma_prefetch_desc(mapA_lo);
    umma::tma_prefetch_desc(&p.tmB_hi);
    umma::tma_prefetch_desc(&p.tmB_lo);
    if (EPI != EPI_PLAIN) umma::tma_prefetch_desc(&p.tmR);
  }
  if (warp == 1) {
    umma::tmem_alloc(umma::smem_u32(&aux->tmem_base), C::TMEM_COLS);
    umma::tmem_relinquish();
  }
  if (EPI != EPI_PLAIN && threadIdx.x >= 64 && threadIdx.x < 72)
    aux->info[threadIdx.x - 64] = __ldg(p.tile_info + (int64_t)tile * 8 + (threadIdx.x - 64));
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  if (mc > 1) umma::cluster_sync_all();           // the peers' barriers exist before anything is multicast to them
  const uint32_t tmem = aux->tmem_base;
  constexpr bool BWD = EPI == EPI_BOND_BWD || EPI == EPI_INIT_BWD;
  const int r_row0 = (EPI == EPI_BOND || BWD) ? tile * TM : (EPI == EPI_READOUT ? aux->info[2] : 0);
  TC_STAMP(1);
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch, tile info) only touches
  // data that no earlier kernel of the forward writes; let the next kernel start its own prologue now
  if (threadIdx.x == 0) umma::grid_dep_launch();

  // ------------------------------------------------------------------ main loop (warp-specialised)
  if (warp == 0) {
    // TMA producer: one elected lane streams A (hi, lo) and B (hi, lo) k-chunks through the ring and
    // prefetches the first column chunk of the epilogue's fp32 operand
    if (lane == 0) umma::grid_dep_wait();          // A / R operands are outputs of the previous kernel
    if (lane == 0 && EPI != EPI_PLAIN && !C::R_ALIAS) {
      const uint32_t rb = umma::smem_u32(&aux->r_full[0]);
      umma::mbar_arrive_expect_tx(rb, C::R_BYTES);
      umma::tma_load_2d(&p.tmR, rb, umma::smem_u32(r0_s), p.r_col0 + n0, r_row0);
    }
    for (int kc = 0; kc < p.num_k; ++kc) {
      const int s = kc % STAGES;
      const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
      if (lane == 0) {
        umma::mbar_wait(umma::smem_u32(&aux->empty[s]), ph ^ 1u);
        const uint32_t full = umma::smem_u32(&aux->full[s]);
        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
        umma::mbar_arrive_expect_tx(full, p.fast ? A_BYTES + B_BYTES : STAGE_BYTES);
        umma::tma_load_2d(mapA_hi, full, st, kc * BK, tile * TM);
        if (!p.fast) umma::tma_load_2d(mapA_lo, full, st + A_BYTES, kc * BK, tile * TM);
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
      __syncwarp();
    }
  } else if (warp == 1) {
    // MMA issuer: one lane issues 3 tcgen05.mma per 16-wide k-step (lo.hi + hi.lo + hi.hi)
    const uint32_t idesc = umma::idesc_f16_f32(TM, n_eff);
    const uint32_t idesc_cat = umma::idesc_f16_f32(TM, 2 * BN);
    for (int kc = 0; kc < p.num_k; ++kc) {
      const int s = kc % STAGES;
      const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
      if (lane == 0) {
        umma::mbar_wait(umma::smem_u32(&aux->full[s]), ph);
        umma::tc_fence_after_sync();
#ifdef CGR_FWD_STAMPS
        if (dbg && kc == 0) dbg[6] = clock64();
        if (dbg && kc == p.num_k - 1) dbg[7] = clock64();
#endif
        const uint32_t st = base + (uint32_t)s * STAGE_BYTES;
        const uint64_t da_hi = umma::smem_desc_k_sw128(st);
        const uint64_t da_lo = umma::smem_desc_k_sw128(st + A_BYTES);
        const uint64_t db_hi = umma::smem_desc_k_sw128(st + 2 * A_BYTES);
        const uint64_t db_lo = umma::smem_desc_k_sw128(st + 2 * A_BYTES + B_BYTES);
        const int k_left = p.k_total - kc * BK;                   // K tail: skip k-steps that are all zero padding
        const int ksteps = k_left >= BK ? BK / 16 : (k_left + 15) / 16;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          if (ks >= ksteps) break;
          const uint64_t adv = (uint64_t)(ks * 32 >> 4);          // 16 fp16 = 32 bytes along K inside the swizzle row
          if (p.fast) {
            umma::mma_f16_ss(tmem, da_hi + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
          } else if (C::CAT) {
            // cols [0,BN) += A_hi B_hi^T, cols [BN,2BN) += A_hi B_lo^T (one instruction), then cols [0,BN) += A_lo B_hi^T
            umma::mma_f16_ss(tmem, da_hi + adv, db_hi + adv, idesc_cat, (kc | ks) ? 1u : 0u);
            umma::mma_f16_ss(tmem, da_lo + adv, db_hi + adv, idesc, 1u);
          } else {
            umma::mma_f16_ss(tmem, da_lo + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
            umma::mma_f16_ss(tmem, da_hi + adv, db_lo + adv, idesc, 1u);
            umma::mma_f16_ss(tmem, da_hi + adv, db_hi + adv, idesc, 1u);
          }
        }
        if (mc > 1) umma::mma_commit_mc(umma::smem_u32(&aux->empty[s]), mc_mask);
        else umma::mma_commit(umma::smem_u32(&aux->empty[s]));   // frees the stage when these MMAs retire
        if (kc == p.num_k - 1) umma::mma_commit(umma::smem_u32(&aux->tmem_full));
      }
      __syncwarp();
    }
  } else if (EPI != EPI_PLAIN) {
    // the other warps stage the tile's index rows into shared memory while the GEMM runs
    const int ebase = aux->info[0], ecount = aux->info[1], abase = aux->info[2], acount = aux->info[3];
    constexpr int X1 = BWD ? 1 : 0;        // backward gathers read the reverse bond of every neighbour
    for (int j = threadIdx.x - 64; j < ecount; j += THREADS - 64) {
      aux->src_l[j] = (uint8_t)(__ldg(p.src + ebase + j) - abase);
      aux->idx_l[j] = (uint8_t)((__ldg(p.in_idx + ebase + j) - ebase) ^ X1);
    }
    for (int v = threadIdx.x - 64; v <= acount; v += THREADS - 64)
      aux->ptr_l[v] = (uint16_t)(__ldg(p.in_ptr + abase + v) - ebase);
    // packed neighbour descriptor of every epilogue row, read straight from the CSR (L2 hits, hidden
    // behind the GEMM): one 8-byte word + degree instead of a chain of dependent shared-memory loads
    const int nrows = (EPI == EPI_BOND || EPI == EPI_BOND_BWD) ? ecount : acount;
    for (int r = threadIdx.x - 64; r < nrows; r += THREADS - 64) {
      // BOND: in-bonds of src(r).  BOND_BWD: in-bonds of dst(r) = src(r^1), each replaced by its reverse.
      // READOUT / INIT_BWD: in-bonds of atom r (INIT_BWD: their reverses = the bonds leaving r).
      const int a = EPI == EPI_BOND ? __ldg(p.src + ebase + r)
                                    : (EPI == EPI_BOND_BWD ? __ldg(p.src + ebase + (r ^ 1)) : abase + r);
      const int pb = __ldg(p.in_ptr + a), pe = __ldg(p.in_ptr + a + 1);
      uint32_t w[2] = {0u, 0u};
      for (int t = 0; t < NBR && pb + t < pe; ++t)
        w[t >> 2] |= (uint32_t)(((__ldg(p.in_idx + pb + t) - ebase) ^ X1) & 0xff) << (8 * (t & 3));
      aux->nbr[r] = make_uint2(w[0], w[1]);
      aux->nbr_pb[r] = (uint16_t)(pb - ebase);
      aux->nbr_deg[r] = (uint8_t)(pe - pb > 255 ? 255 : pe - pb);
    }
  }

  const float us_w = __ldg(p.unscale);
  const float skip = ((EPI == EPI_BOND || EPI == EPI_BOND_BWD) && p.skip) ? __ldg(p.skip) : 1.f;
  const int c = 4 * lane;                                         // this lane's float4 column group inside a chunk
  // per-lane bias / readout weights of every column chunk, fetched while the GEMM is still running
  float4 bias_r[NCH], wf_r[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    bias_r[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
    wf_r[ch] = bias_r[ch];
    const int n = n0 + ch * CH + c;
    if (lane < VL && n < p.n_total) {
      if (p.bias) {
        bias_r[ch].x = __ldg(p.bias + n);
        bias_r[ch].y = n + 1 < p.n_total ? __ldg(p.bias + n + 1) : 0.f;
        bias_r[ch].z = n + 2 < p.n_total ? __ldg(p.bias + n + 2) : 0.f;
        bias_r[ch].w = n + 3 < p.n_total ? __ldg(p.bias + n + 3) : 0.f;
      }
      if (EPI == EPI_READOUT) {
        wf_r[ch].x = __ldg(p.w_ffn + n);
        wf_r[ch].y = n + 1 < p.n_total ? __ldg(p.w_ffn + n + 1) : 0.f;
        wf_r[ch].z = n + 2 < p.n_total ? __ldg(p.w_ffn + n + 2) : 0.f;
        wf_r[ch].w = n + 3 < p.n_total ? __ldg(p.w_ffn + n + 3) : 0.f;
      }
    }
  }

  // ------------------------------------------------------------------ epilogue (all 8 warps)
  umma::mbar_wait(umma::smem_u32(&aux->tmem_full), 0);
  umma::tc_fence_after_sync();
  // backward: the A operand carries the power-of-two scale fixed by the amax its producer saw; the previous kernel
  // has completed (the MMAs consumed its output), so these scalars are final
  float us = us_w;
  float out_scale = 1.f;
  if (BWD) {
    us /= grad_scale(__ldcg(p.gamax_in));
    out_scale = grad_scale(__ldcg(p.gamax_out));
    if (p.gunscale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.gunscale_out = 1.f / out_scale;
  }
  TC_STAMP(2);
  if (C::R_ALIAS && EPI != EPI_PLAIN && threadIdx.x == 0) {
    // every MMA has retired and every pipeline load has landed: the stage buffers are free for the fp32 operand
    const uint32_t rb = umma::smem_u32(&aux->r_full[0]);
    umma::mbar_arrive_expect_tx(rb, C::R_BYTES);
    umma::tma_load_2d(&p.tmR, rb, umma::smem_u32(r0_s), p.r_col0 + n0, r_row0);
  }
  if (NCH > 1 && EPI != EPI_PLAIN && threadIdx.x == 0) {
    // second column chunk of the fp32 operand goes into the drained pipeline buffers
    const uint32_t rb = umma::smem_u32(&aux->r_full[1]);
    umma::mbar_arrive_expect_tx(rb, C::R_BYTES);
    umma::tma_load_2d(&p.tmR, rb, umma::smem_u32(r1_s), p.r_col0 + n0 + CH, r_row0);
  }
  if (EPI == EPI_READOUT) {
    for (int v = threadIdx.x; v < TM; v += THREADS) aux->tat[v] = 0.f;
  }
  bool ovf = false;

#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (ch > 0) __syncthreads();                                  // every reader of the previous y_s chunk is done
    // TMEM -> registers -> y_s (fp32, unscaled); warp w owns lane quarter w%4 and every other 8-column group
    {
      const int q = warp & 3, grp = warp >> 2;
      const int row = q * 32 + lane;
#pragma unroll
      for (int cc = grp * 8; cc < CH; cc += 8 * (NWARPS / 4)) {
        float v[8];
        umma::tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * CH + cc), v);
        if (C::CAT && !p.fast) {
          float v2[8];
          umma::tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + ch * CH + cc), v2);
          umma::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += v2[i];
        } else {
          umma::tmem_ld_wait();
        }
        float4* dst = reinterpret_cast<float4*>(y_s + row * CHP + cc);
        dst[0] = make_float4(v[0] * us, v[1] * us, v[2] * us, v[3] * us);
        dst[1] = make_float4(v[4] * us, v[5] * us, v[6] * us, v[7] * us);
      }
    }
    umma::tc_fence_before_sync();
    if (EPI != EPI_PLAIN) umma::mbar_wait(umma::smem_u32(&aux->r_full[ch]), 0);   // fp32 operand chunk has landed
    __syncthreads();
    if (ch == 0) TC_STAMP(3);

    const int n = n0 + ch * CH + c;                               // first global column of this lane's group
    const bool lane_on = lane < VL && n < p.n_total;
    const float4 bias4 = bias_r[NCH > 1 ? ch : 0], wf4 = wf_r[NCH > 1 ? ch : 0];
    const float* r_s = ch == 0 ? r0_s : r1_s;

    if (EPI == EPI_PLAIN) {
      // rows are dense (atoms): out = y + bias, written as coalesced rows
      if (lane_on) {
        for (int r = warp; r < TM; r += NWARPS) {
          const int64_t row = (int64_t)tile * TM + r;
          if (row >= m_rows) break;
          const float4 y = ld4(y_s + r * CHP + c);
          float* o = out_f32 + row * p.ldc + n;
          if (n + 3 < p.n_total && (p.ldc & 3) == 0) {
            *reinterpret_cast<float4*>(o) = make_float4(y.x + bias4.x, y.y + bias4.y, y.z + bias4.z, y.w + bias4.w);
          } else {
            o[0] = y.x + bias4.x;
            if (n + 1 < p.n_total) o[1] = y.y + bias4.y;
            if (n + 2 < p.n_total) o[2] = y.z + bias4.z;
            if (n + 3 < p.n_total) o[3] = y.w + bias4.w;
          }
        }
      }
    } else if (EPI == EPI_BOND) {
      // z[e] = sum_{k in in(src e)} y[k] - y[e^1] + b + skip*h0[e];  h' = dropout(act(z)); RU rows per warp step
      const int ebase = aux->info[0], ecount = aux->info[1];
      const int H = p.n_total;
      const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
      float vmax = 0.f;
      for (int j0 = warp * RU; j0 < ecount; j0 += NWARPS * RU) {
        uint2 nb[RU];
        int deg[RU], maxdeg = 0;
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const int j = j0 + u < ecount ? j0 + u : j0;
          nb[u] = aux->nbr[j];
          deg[u] = j0 + u < ecount ? (int)aux->nbr_deg[j] : 0;
          maxdeg = deg[u] > maxdeg ? deg[u] : maxdeg;
        }
        if (lane_on) {
          float4 acc[RU];
#pragma unroll
          for (int u = 0; u < RU; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          // ascending bond id: the reference's order.  Slots 0-3 are unrolled with constant shifts (a CGR atom
          // rarely has more than 4 bonds), slots 4-7 loop.
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (t < maxdeg) {
#pragma unroll
              for (int u = 0; u < RU; ++u)
                if (t < deg[u]) add4(acc[u], ld4(y_s + (int)((nb[u].x >> (8 * t)) & 0xffu) * CHP + c));
            }
          }
          const int fast = maxdeg < NBR ? maxdeg : NBR;
          for (int t = 4; t < fast; ++t) {
#pragma unroll
            for (int u = 0; u < RU; ++u)
              if (t < deg[u]) add4(acc[u], ld4(y_s + (int)((nb[u].y >> (8 * (t - 4))) & 0xffu) * CHP + c));
          }
          if (maxdeg > NBR) {                                     // rare: atoms with more than NBR bonds
#pragma unroll
            for (int u = 0; u < RU; ++u) {
              if (j0 + u >= ecount) break;
              const int pb = aux->nbr_pb[j0 + u];
              for (int t = NBR; t < deg[u]; ++t) add4(acc[u], ld4(y_s + (int)aux->idx_l[pb + t] * CHP + c));
            }
          }
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int j = j0 + u;
            if (j >= ecount) break;
            const int64_t r = (int64_t)tile * TM + j;
            const float4 yr = ld4(y_s + (j ^ 1) * CHP + c);
            const float4 h0v = ld4(r_s + j * CH + c);
            float z[4] = {acc[u].x - yr.x + bias4.x + skip * h0v.x, acc[u].y - yr.y + bias4.y + skip * h0v.y,
                          acc[u].z - yr.z + bias4.z + skip * h0v.z, acc[u].w - yr.w + bias4.w + skip * h0v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float v = act_t<RELU>(z[i], p.act);
              if (p.dropout_p > 0.f)
                v = cgr_dropout_keep(p.seed, p.layer, (uint64_t)(ebase + j) * (uint64_t)H + (uint64_t)(n + i),
                                     p.dropout_p) ? v * keep_scale : 0.f;
              z[i] = v;                                            // H % 4 == 0: the whole column group is real
            }
            vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(z[0]), fabsf(z[1]))), fmaxf(fabsf(z[2]), fabsf(z[3])));
            // FP16 (hi, lo) split with packed conversions: hi = rn(v), lo = rn(v - hi)
            const __half2 hi01 = __floats2half2_rn(z[0], z[1]), hi23 = __floats2half2_rn(z[2], z[3]);
            const float2 f01 = __half22float2(hi01), f23 = __half22float2(hi23);
            const __half2 lo01 = __floats2half2_rn(z[0] - f01.x, z[1] - f01.y);
            const __half2 lo23 = __floats2half2_rn(z[2] - f23.x, z[3] - f23.y);
            uint2 ph, pl;
            ph.x = *reinterpret_cast<const uint32_t*>(&hi01);
            ph.y = *reinterpret_cast<const uint32_t*>(&hi23);
            pl.x = *reinterpret_cast<const uint32_t*>(&lo01);
            pl.y = *reinterpret_cast<const uint32_t*>(&lo23);
            *reinterpret_cast<uint2*>(p.o_hi + r * p.ldo + n) = ph;
            *reinterpret_cast<uint2*>(p.o_lo + r * p.ldo + n) = pl;
          }
        }
      }
      ovf |= vmax > 60000.f;
    } else if (BWD) {
      // y_s holds dh (BOND_BWD: dL/dh_l of this tile's bonds; INIT_BWD: the layer-1 part of dL/dh_0).
      //   pass 1 (row-local): dz = (dh [+ dh0_acc]) * [h > 0] * keep_scale, written back to y_s; column sums of dz
      //           (bias gradient), sum of dz . h0 (skip-weight gradient), dh0_acc += skip * dz.
      //   pass 2 (tile-local gather): BOND_BWD  dy[k] = sum_{j in in(dst k)} dz[j^1] - dz[k^1]   (transpose of the
      //           forward gather);  INIT_BWD  dP[v] = sum_{j in in(v)} dz[j^1]  (bonds leaving atom v).
      const int ebase = aux->info[0], ecount = aux->info[1], abase = aux->info[2], acount = aux->info[3];
      const int H = p.n_total;
      float* red_s = reinterpret_cast<float*>(smem + C::RED_OFF);
      float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
      float dsk = 0.f;
      if (lane_on) {
        // PU rows per step: their global loads (mask, dh0_acc) are issued together, then consumed
        constexpr int PU = 4;
        for (int j0 = warp; j0 < ecount; j0 += NWARPS * PU) {
          uint2 mh[PU];
          float4 a0[PU];
#pragma unroll
          for (int u = 0; u < PU; ++u) {
            const int j = j0 + u * NWARPS;
            if (j < ecount) {
              const int64_t r = (int64_t)tile * TM + j;
              mh[u] = __ldg(reinterpret_cast<const uint2*>(p.mask_hi + r * p.ld_mask + n));
              if (EPI == EPI_BOND_BWD && !p.dh0_first)
                a0[u] = __ldcg(reinterpret_cast<const float4*>(p.dh0_acc + r * H + n));   // previous kernel's: bypass L1
              else
                a0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < PU; ++u) {
            const int j = j0 + u * NWARPS;
            if (j >= ecount) break;
            const int64_t r = (int64_t)tile * TM + j;
            float4 dh = ld4(y_s + j * CHP + c);
            if (EPI == EPI_INIT_BWD) add4(dh, ld4(r_s + j * CH + c));
            const __half2 m01 = *reinterpret_cast<const __half2*>(&mh[u].x), m23 = *reinterpret_cast<const __half2*>(&mh[u].y);
            const float2 f01 = __half22float2(m01), f23 = __half22float2(m23);
            float4 dz;
            dz.x = f01.x > 0.f ? dh.x * p.keep_scale : 0.f;
            dz.y = f01.y > 0.f ? dh.y * p.keep_scale : 0.f;
            dz.z = f23.x > 0.f ? dh.z * p.keep_scale : 0.f;
            dz.w = f23.y > 0.f ? dh.w * p.keep_scale : 0.f;
            *reinterpret_cast<float4*>(y_s + j * CHP + c) = dz;
            add4(csum, dz);
            if (EPI == EPI_BOND_BWD) {
              if (p.dskip_partial) {
                const float4 h0v = ld4(r_s + j * CH + c);
                dsk = fmaf(dz.x, h0v.x, fmaf(dz.y, h0v.y, fmaf(dz.z, h0v.z, fmaf(dz.w, h0v.w, dsk))));
              }
              float4 acc = a0[u];
              acc.x = fmaf(skip, dz.x, acc.x); acc.y = fmaf(skip, dz.y, acc.y);
              acc.z = fmaf(skip, dz.z, acc.z); acc.w = fmaf(skip, dz.w, acc.w);
              *reinterpret_cast<float4*>(p.dh0_acc + r * H + n) = acc;
            } else {
              *reinterpret_cast<float4*>(p.dz0_out + (int64_t)(ebase + j) * H + n) = dz;
            }
          }
        }
      }
      if (lane < VL) *reinterpret_cast<float4*>(red_s + warp * CH + c) = csum;
      if (EPI == EPI_BOND_BWD && p.dskip_partial) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dsk += __shfl_xor_sync(0xffffffffu, dsk, o);
        if (lane == 0) aux->red2[warp] = dsk;
      }
      __syncthreads();
      if (threadIdx.x < CH) {                                      // fixed warp order: deterministic
        const int col = n0 + ch * CH + threadIdx.x;
        if (col < p.n_total) {
          float sres = 0.f;
          for (int w = 0; w < NWARPS; ++w) sres += red_s[w * CH + threadIdx.x];
          p.colsum_partial[(int64_t)tile * H + col] = sres;
        }
      }
      if (EPI == EPI_BOND_BWD && p.dskip_partial && threadIdx.x == 0) {
        float sres = 0.f;
        for (int w = 0; w < NWARPS; ++w) sres += aux->red2[w];
        p.dskip_partial[((int64_t)tile * gridDim.x + slice) * NCH + ch] = sres;
      }
      float vmax = 0.f;
      const int nrows = EPI == EPI_BOND_BWD ? ecount : acount;
      for (int j0 = warp * RU; j0 < nrows; j0 += NWARPS * RU) {
        float4 acc[RU];
        if (lane_on) {
          nbr_sum<CHP>(y_s, aux, j0, nrows, c, acc);
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int j = j0 + u;
            if (j >= nrows) break;
            float4 g = acc[u];
            int64_t orow;
            if (EPI == EPI_BOND_BWD) {
              const float4 yr = ld4(y_s + (j ^ 1) * CHP + c);
              g.x -= yr.x; g.y -= yr.y; g.z -= yr.z; g.w -= yr.w;
              orow = (int64_t)tile * TM + j;
            } else {
              orow = abase + j;
            }
            vmax = fmaxf(vmax, amax4(g));
            store_split4(g, out_scale, p.o_hi + orow * p.ldo + n, p.o_lo + orow * p.ldo + n);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (lane == 0 && vmax > 0.f) atomicMax(p.gamax_track, __float_as_uint(vmax));   // max is order-independent
      ovf |= vmax * out_scale > 60000.f;
    } else {
      // readout: hv[v] = act(Q[v] + sum_{k in in(v)} y[k]);  t[v] += hv[v] . w_f over this chunk's columns
      const int acount = aux->info[3];
      for (int v0 = warp * RU; v0 < acount; v0 += NWARPS * RU) {
        uint2 nb[RU];
        int deg[RU], maxdeg = 0;
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const int v = v0 + u < acount ? v0 + u : v0;
          nb[u] = aux->nbr[v];
          deg[u] = v0 + u < acount ? (int)aux->nbr_deg[v] : 0;
          maxdeg = deg[u] > maxdeg ? deg[u] : maxdeg;
        }
        float t[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) t[u] = 0.f;
        if (lane_on) {
          float4 acc[RU];
#pragma unroll
          for (int u = 0; u < RU; ++u) acc[u] = v0 + u < acount ? ld4(r_s + (v0 + u) * CH + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < maxdeg) {
#pragma unroll
              for (int u = 0; u < RU; ++u)
                if (k < deg[u]) add4(acc[u], ld4(y_s + (int)((nb[u].x >> (8 * k)) & 0xffu) * CHP + c));
            }
          }
          const int fast = maxdeg < NBR ? maxdeg : NBR;
          for (int k = 4; k < fast; ++k) {
#pragma unroll
            for (int u = 0; u < RU; ++u)
              if (k < deg[u]) add4(acc[u], ld4(y_s + (int)((nb[u].y >> (8 * (k - 4))) & 0xffu) * CHP + c));
          }
          if (maxdeg > NBR) {
#pragma unroll
            for (int u = 0; u < RU; ++u) {
              if (v0 + u >= acount) break;
              const int pb = aux->nbr_pb[v0 + u];
              for (int k = NBR; k < deg[u]; ++k) add4(acc[u], ld4(y_s + (int)aux->idx_l[pb + k] * CHP + c));
            }
          }
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const float4 hv = make_float4(act_t<RELU>(acc[u].x, p.act), act_t<RELU>(acc[u].y, p.act),
                                          act_t<RELU>(acc[u].z, p.act), act_t<RELU>(acc[u].w, p.act));
            t[u] = hv.x * wf4.x;
            t[u] = fmaf(hv.y, wf4.y, t[u]);
            t[u] = fmaf(hv.z, wf4.z, t[u]);
            t[u] = fmaf(hv.w, wf4.w, t[u]);
            if (p.hv_out && v0 + u < acount)
              *reinterpret_cast<float4*>(p.hv_out + (int64_t)(aux->info[2] + v0 + u) * p.n_total + n) = hv;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int u = 0; u < RU; ++u) t[u] += __shfl_xor_sync(0xffffffffu, t[u], o);
        }
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < RU; ++u)
            if (v0 + u < acount) aux->tat[v0 + u] += t[u];       // the same warp owns atom v in every chunk
        }
      }
    }
  }

  if ((EPI == EPI_BOND || BWD) && ovf) atomicOr(p.overflow, 1);
  if (EPI == EPI_READOUT) {
    __syncthreads();
    const int abase = aux->info[2], rx0 = aux->info[4], rxcount = aux->info[5];
    for (int rx = threadIdx.x; rx < rxcount; rx += THREADS) {
      const int b = rx0 + rx;
      const int v0 = __ldg(p.atom_ptr + b) - abase, v1 = __ldg(p.atom_ptr + b + 1) - abase;
      float s = 0.f;
      for (int v = v0; v < v1; ++v) s += aux->tat[v];            // ascending atom id
      p.partial_out[(int64_t)slice * p.n_rxn + b] = s;
    }
    // the last slice CTA of this tile to arrive adds the slices in a fixed order: deterministic, no extra kernel
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) aux->info[7] = atomicAdd(p.tile_counter + tile, 1);
    __syncthreads();
    if (aux->info[7] == (int)gridDim.x - 1) {
      __threadfence();
      const float bf = __ldg(p.b_ffn);
      // an operand left the fp16 range somewhere in this forward (the flag is final: every producer kernel has completed):
      // the energies are poisoned with NaN so the condition cannot go unnoticed without a host synchronisation
      const bool poisoned = p.overflow && (__ldcg(p.overflow) & 3) != 0;
      for (int rx = threadIdx.x; rx < rxcount; rx += THREADS) {
        const int b = rx0 + rx;
        float s = 0.f;
        for (int i = 0; i < (int)gridDim.x; ++i) s += __ldcg(p.partial_out + (int64_t)i * p.n_rxn + b);
        p.out[b] = poisoned ? __int_as_float(0x7fc00000) : s + bf;
      }
      if (threadIdx.x == 0) p.tile_counter[tile] = 0;             // ready for the next forward
    }
  }

  __syncthreads();
  if (mc > 1) umma::cluster_sync_all();           // no CTA leaves while a peer may still arrive on its barriers
  TC_STAMP(5);
  if (warp == 1) umma::tmem_dealloc(tmem, C::TMEM_COLS);
#undef TC_STAMP
}

}  // namespace tcg
