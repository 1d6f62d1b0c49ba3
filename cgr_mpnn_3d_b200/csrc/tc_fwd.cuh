// Fused message-passing forward of the tcgen05 engine: ALL bond layers + the readout of a group of row tiles in ONE
// launch (reference cgr_mpnn_3D/models/GNN.py:90-110).
//
// Every stage of the path is tile-local because a 128-row tile holds whole reactions, so a tile never needs data of
// another tile: a thread-block CLUSTER of S CTAs (one per BN-wide column slice of the hidden dimension) takes a group
// of one or two tiles through every layer.  Per layer a CTA computes its slice  y = h W_l^T  (TMA-fed FP16x3
// tcgen05.mma, accumulator in TMEM), gathers  z[e] = sum_{k in in(src e)} y[k] - y[e^1] + b + skip*h0[e]  in shared
// memory, applies the activation and writes its slice of the next operand (hi, lo); the S slices of a tile meet
// through an mbarrier in every CTA of the cluster that the peers arrive on remotely (release / acquire at cluster
// scope, plus a proxy fence because the next layer reads the operand with TMA).  The cluster is gang-scheduled, so the
// wait can never deadlock, whatever else runs on the device.
//
// Warp roles (persistent over the items (tile, layer) of the group):
//   warp 0       TMA producer: weight chunks are requested as soon as a stage is free (they never depend on a peer),
//                the activation chunks once the tile's previous layer is complete cluster-wide
//   warp 1       tcgen05.mma issuer, accumulators double-buffered in TMEM (2 x 256 columns)
//   warp 2       TMEM allocation
//   warps 4..15  epilogue: TMEM -> fp32 staging rows -> gather / activation / split -> global
// With two tiles per cluster the epilogue of tile X, layer l overlaps the MMAs of tile Y, layer l, and so on:
// the dependency chain of one tile hides behind the other tile's work.
#pragma once
#include "tc_gemm.cuh"

namespace tcf {

using tcg::A_BYTES;
using tcg::BK;
using tcg::NBR;
using tcg::TM;

constexpr int MAX_LAYERS = 16;                 // bond layers + readout
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS;     // 640
constexpr int ZROW = TM;                       // staging row that stays zero: target of unused neighbour slots
constexpr int SLOT_COLS = 256;                 // TMEM columns per accumulator slot
constexpr int MAX_TPC = 2;                     // tiles per cluster (ping-pong)
constexpr int SMEM_LIMIT = 232448;

template <int BN_>
struct FCfg {
  static constexpr int BN = BN_;
  static constexpr bool CAT = 2 * BN <= 256;                   // A_hi x [B_hi ; B_lo] as one MMA of N = 2 BN
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int NCH = BN > 128 ? 4 : 1;                 // epilogue column chunks
  static constexpr int CH = BN / NCH;                          // 52 (BN = 208), 80, 112
  static constexpr int UPR = CH / 4;                           // float4 units per row of a chunk
  static constexpr int CH8 = (CH + 7) / 8 * 8;
  static constexpr int CHP = ((CH8 / 4) % 2 == 1) ? CH8 : CH8 + 4;   // staging row pitch: odd multiple of 4 floats
  static constexpr int Y_BYTES = ((TM + 1) * CHP * 4 + 1023) / 1024 * 1024;     // + the zero row
  static constexpr int LPR = UPR <= 16 ? 16 : 32;              // lanes per row in the readout epilogue
  static constexpr int AUX_BYTES = 12288;
  static constexpr int FIT = (SMEM_LIMIT - 1024 - Y_BYTES - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 4 ? 4 : FIT;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + Y_BYTES + AUX_BYTES;
  static_assert(BN % 16 == 0 && BN <= 256 && BN % NCH == 0 && CH % 4 == 0 && UPR <= 32, "bad slice width");
  static_assert(STAGES >= 2, "pipeline needs two stages");
  static_assert((CAT ? 2 * BN : BN) <= SLOT_COLS, "accumulator exceeds its TMEM slot");
};

struct FwdParams {
  CUtensorMap tmA_hi[2], tmA_lo[2];            // activation operand, ping-pong: layer l reads buffer l & 1
  CUtensorMap tmB_hi[MAX_LAYERS], tmB_lo[MAX_LAYERS];   // prepared weights of bond layer l; [depth] = W_os (readout)
  const float* bias[MAX_LAYERS];               // bond-layer biases
  const float* skip[MAX_LAYERS];               // learnable skip scalars (device) or null (= 1)
  __half* o_hi[2];                             // layer l writes buffer (l + 1) & 1, rows tile * 128 + j
  __half* o_lo[2];
  int64_t ldo;
  const float* unscale;                        // [1 + l]: 1 / weight scale of layer l's matrix
  const float* h0;                             // [T * 128, H] fp32, tile-packed (skip operand)
  const float* PQ;                             // [N, 2H] fp32: Q' = PQ[:, H:] (readout operand)
  const float* w_ffn;
  const float* b_ffn;
  const int32_t* tile_info;
  const int32_t* in_ptr;
  const int32_t* in_idx;
  const int32_t* src;
  const int32_t* atom_ptr;
  float* partial_out;                          // [S, B]
  float* out;                                  // [B]
  int* tile_counter;                           // [T]: low 16 bits arrival counter, bit 16 = the tile overflowed
  int* overflow;                               // tc_status[0]
  int64_t n_rxn;
  int depth, H, num_k, act, n_tiles, tiles_per_cluster;
  int fast;                                    // 1: single-pass fp16 (hi halves only), the "fast" precision mode
  long long* dbg;                              // optional [n_cta][MAX_TPC * MAX_LAYERS][4] clock64 stamps (debug)
};

struct TileAux {                // per tile of the group: packed neighbour descriptors (built once, used by every layer)
  int32_t info[8];
  // bond row j: the in-bonds of src(j) EXCEPT the reverse bond j^1 (GNN.py:141 adds it and subtracts it again), atom row
  // v: the in-bonds of v -- tile-local row ids, one byte each, ascending bond id, unused slots = ZROW
  uint2 nbr_b[TM];
  uint2 nbr_a[TM];
  uint16_t pb_b[TM];            // CSR offset of the row's full in-bond list (rows with more than NBR neighbours)
  uint16_t pb_a[TM];
  uint8_t cnt_b[TM];            // neighbours of the row (bond rows: reverse excluded)
  uint8_t cnt_a[TM];
  uint8_t full_b[TM];           // length of the full CSR list behind pb_b
  uint8_t idx_l[TM];            // tile-local bond ids grouped by target atom
  float tat[TM];                // readout: per-atom dot with w_ffn
  int32_t ticket;
};
struct Aux {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t ready[MAX_TPC];      // tile j's operand of the next layer is complete in every CTA of the cluster
  uint32_t tmem_base;
  alignas(16) float bias_s[2][256];   // bias slice of the item being drained (double-buffered by accumulator slot)
  TileAux t[MAX_TPC];
};
static_assert(sizeof(Aux) <= 12288, "Aux too large");

// z -> activation -> FP16 (hi, lo) of one float4 unit, stored as two 8-byte words; returns max |h|
template <bool RELU>
__device__ __forceinline__ float act_split_store(float4 z, int act, __half* hi, __half* lo) {
  z.x = tcg::act_t<RELU>(z.x, act); z.y = tcg::act_t<RELU>(z.y, act);
  z.z = tcg::act_t<RELU>(z.z, act); z.w = tcg::act_t<RELU>(z.w, act);
  tcg::store_split4(z, 1.f, hi, lo);
  return tcg::amax4(z);
}

// four consecutive floats of a parameter vector (16-byte aligned in practice; parameters may also be views)
__device__ __forceinline__ float4 ldg4(const float* p) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// sum of the staged rows a packed descriptor names: slots 0..3 unconditionally (unused slots point at the zero row: no
// branches, four independent loads), slots 4..7 only for rows with more than four neighbours
template <int CHP>
__device__ __forceinline__ float4 gather_packed(const float* y_c, uint2 nb, int cnt) {
  float4 a = tcg::ld4(y_c + (int)(nb.x & 0xffu) * CHP);
  const float4 v1 = tcg::ld4(y_c + (int)((nb.x >> 8) & 0xffu) * CHP);
  const float4 v2 = tcg::ld4(y_c + (int)((nb.x >> 16) & 0xffu) * CHP);
  const float4 v3 = tcg::ld4(y_c + (int)(nb.x >> 24) * CHP);
  tcg::add4(a, v1); tcg::add4(a, v2); tcg::add4(a, v3);
  if (cnt > 4) {
    const float4 v4 = tcg::ld4(y_c + (int)(nb.y & 0xffu) * CHP);
    const float4 v5 = tcg::ld4(y_c + (int)((nb.y >> 8) & 0xffu) * CHP);
    const float4 v6 = tcg::ld4(y_c + (int)((nb.y >> 16) & 0xffu) * CHP);
    const float4 v7 = tcg::ld4(y_c + (int)(nb.y >> 24) * CHP);
    tcg::add4(a, v4); tcg::add4(a, v5); tcg::add4(a, v6); tcg::add4(a, v7);
  }
  return a;
}
// rows with more than NBR neighbours (rare: an atom with ten or more bonds): walk the CSR list, skipping row `skip`
template <int CHP>
__device__ __forceinline__ float4 gather_list(const float* y_c, const uint8_t* idx_l, int pb, int n, int skip) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < n; ++t) {
    const int k = idx_l[pb + t];
    if (k != skip) tcg::add4(a, tcg::ld4(y_c + k * CHP));
  }
  return a;
}

template <int BN_, bool RELU>
__global__ void __launch_bounds__(THREADS, 1) tc_fwd_kernel(const __grid_constant__ FwdParams p) {
  using C = FCfg<BN_>;
  constexpr int BN = C::BN, CH = C::CH, CHP = C::CHP, UPR = C::UPR, NCH = C::NCH, STAGES = C::STAGES;
  constexpr int STAGE_BYTES = C::STAGE_BYTES, B_BYTES = C::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = umma::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  float* y_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  Aux* aux = reinterpret_cast<Aux*>(smem + STAGES * STAGE_BYTES + C::Y_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = (int)umma::cluster_nctarank();
  const int slice = (int)umma::cluster_ctarank();
  const int group = (int)blockIdx.x / S;
  const int n0 = slice * BN;
  const int H = p.H, depth = p.depth;
  int n_eff = H - n0;                                              // columns this slice owns, rounded to the MMA granularity
  n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
  const int tile0 = group * p.tiles_per_cluster;
  int nt = p.n_tiles - tile0;
  nt = nt < p.tiles_per_cluster ? nt : p.tiles_per_cluster;
  const int n_items = nt * (depth + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->tmem_full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->tmem_empty[s]), EPI_WARPS);
    }
    for (int j = 0; j < MAX_TPC; ++j) umma::mbar_init(umma::smem_u32(&aux->ready[j]), (uint32_t)S);
    umma::mbar_fence_init();
    umma::tma_prefetch_desc(&p.tmA_hi[0]);
    umma::tma_prefetch_desc(&p.tmA_lo[0]);
    umma::tma_prefetch_desc(&p.tmB_hi[0]);
    umma::tma_prefetch_desc(&p.tmB_lo[0]);
  }
  if (warp == 2) {
    umma::tmem_alloc(umma::smem_u32(&aux->tmem_base), 512);
    umma::tmem_relinquish();
  }
  if (threadIdx.x >= 96 && threadIdx.x < 96 + 8 * MAX_TPC) {
    const int j = (threadIdx.x - 96) >> 3, k = (threadIdx.x - 96) & 7;
    aux->t[j].info[k] = j < nt ? __ldg(p.tile_info + (int64_t)(tile0 + j) * 8 + k) : 0;
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  // every CTA's barriers exist before a peer may arrive on them
  umma::cluster_sync_all();
  const uint32_t tmem = aux->tmem_base;
  // programmatic dependent launch: nothing above touches data an earlier kernel of the forward writes
  if (threadIdx.x == 0) umma::grid_dep_launch();

  // ---- packed neighbour descriptors of the group's tiles (index arrays are batch preparation: no dependency) ----
  for (int j = 0; j < nt; ++j) {
    TileAux& ta = aux->t[j];
    const int ebase = ta.info[0], ecount = ta.info[1], abase = ta.info[2], acount = ta.info[3];
    for (int i = threadIdx.x; i < ecount; i += THREADS) ta.idx_l[i] = (uint8_t)(__ldg(p.in_idx + ebase + i) - ebase);
    for (int r = threadIdx.x; r < ecount + acount; r += THREADS) {
      const bool bond = r < ecount;
      const int a = bond ? __ldg(p.src + ebase + r) : abase + (r - ecount);
      const int skip = bond ? (r ^ 1) : -1;
      const int pb = __ldg(p.in_ptr + a), pe = __ldg(p.in_ptr + a + 1);
      uint32_t w[2] = {0x80808080u, 0x80808080u};             // ZROW in every slot
      int cnt = 0;
      for (int t = pb; t < pe; ++t) {
        const int k = __ldg(p.in_idx + t) - ebase;
        if (k == skip) continue;
        if (cnt < NBR) w[cnt >> 2] = (w[cnt >> 2] & ~(0xffu << (8 * (cnt & 3)))) | ((uint32_t)(k & 0xff) << (8 * (cnt & 3)));
        ++cnt;
      }
      if (bond) {
        ta.nbr_b[r] = make_uint2(w[0], w[1]); ta.pb_b[r] = (uint16_t)(pb - ebase);
        ta.cnt_b[r] = (uint8_t)(cnt > 255 ? 255 : cnt); ta.full_b[r] = (uint8_t)(pe - pb > 255 ? 255 : pe - pb);
      } else {
        const int v = r - ecount;
        ta.nbr_a[v] = make_uint2(w[0], w[1]); ta.pb_a[v] = (uint16_t)(pb - ebase); ta.cnt_a[v] = (uint8_t)(cnt > 255 ? 255 : cnt);
      }
    }
    for (int v = threadIdx.x; v < TM; v += THREADS) ta.tat[v] = 0.f;
  }
  for (int k = threadIdx.x; k < CHP; k += THREADS) y_s[ZROW * CHP + k] = 0.f;
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      umma::grid_dep_wait();                     // h_0 (layer 0's operand), PQ: outputs of the previous kernels
      uint32_t g = 0;                            // k-chunks issued so far (ring position)
      for (int i = 0; i < n_items; ++i) {
        const int j = i % nt, l = i / nt;
        const int tile = tile0 + j;
        const int buf = l & 1;
        for (int kc = 0; kc < p.num_k; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          umma::mbar_wait(umma::smem_u32(&aux->empty[s]), ph ^ 1u);
          const uint32_t full = umma::smem_u32(&aux->full[s]);
          const uint32_t st = base + s * STAGE_BYTES;
          umma::mbar_arrive_expect_tx(full, p.fast ? A_BYTES + B_BYTES : STAGE_BYTES);
          umma::tma_load_2d(&p.tmB_hi[l], full, st + 2 * A_BYTES, kc * BK, n0);
          if (!p.fast) umma::tma_load_2d(&p.tmB_lo[l], full, st + 2 * A_BYTES + B_BYTES, kc * BK, n0);
          if (kc == 0 && l > 0) {
            // the S slices of this tile's previous layer have been stored (every CTA of the cluster arrived)
            umma::mbar_wait_cluster(umma::smem_u32(&aux->ready[j]), (uint32_t)(l - 1) & 1u);
            umma::fence_proxy_async();
            if (p.dbg) p.dbg[(int64_t)blockIdx.x * (MAX_TPC * MAX_LAYERS * 4) + i * 4 + 3] = clock64();
          }
          umma::tma_load_2d(&p.tmA_hi[buf], full, st, kc * BK, tile * TM);
          if (!p.fast) umma::tma_load_2d(&p.tmA_lo[buf], full, st + A_BYTES, kc * BK, tile * TM);
        }
      }
    }
    __syncwarp();                                // lanes 1..31 wait here for the elected lane
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma::idesc_f16_f32(TM, n_eff);
      const uint32_t idesc_cat = umma::idesc_f16_f32(TM, 2 * BN);
      uint32_t g = 0;
      for (int i = 0; i < n_items; ++i) {
        const uint32_t slot = (uint32_t)i & 1u;
        umma::mbar_wait(umma::smem_u32(&aux->tmem_empty[slot]), (((uint32_t)i >> 1) & 1u) ^ 1u);
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
          const int k_left = H - kc * BK;                          // K tail: skip k-steps that are all zero padding
          const int ksteps = k_left >= BK ? BK / 16 : (k_left + 15) / 16;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            if (ks >= ksteps) break;
            const uint64_t adv = (uint64_t)(ks * 32 >> 4);
            if (p.fast) {
              umma::mma_f16_ss(acc, da_hi + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
            } else if (C::CAT) {
              umma::mma_f16_ss(acc, da_hi + adv, db_hi + adv, idesc_cat, (kc | ks) ? 1u : 0u);
              umma::mma_f16_ss(acc, da_lo + adv, db_hi + adv, idesc, 1u);
            } else {
              umma::mma_f16_ss(acc, da_lo + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
              umma::mma_f16_ss(acc, da_hi + adv, db_lo + adv, idesc, 1u);
              umma::mma_f16_ss(acc, da_hi + adv, db_hi + adv, idesc, 1u);
            }
          }
          umma::mma_commit(umma::smem_u32(&aux->empty[s]));
          if (kc == p.num_k - 1) umma::mma_commit(umma::smem_u32(&aux->tmem_full[slot]));
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 4, et = (int)threadIdx.x - 128;
    const int q = warp & 3, grp = ew >> 2;                          // TMEM lane quarter of this warp, column-group phase
    // bond layers: flat unit mapping -- unit u = (row u / UPR, float4 column group u % UPR); thread et owns the units
    // et, et + EPI_THREADS, ... of every chunk (SLOTS of them)
    constexpr int SLOTS = (TM * UPR + EPI_THREADS - 1) / EPI_THREADS;
    // readout: LPR lanes per atom row; the same lanes own atom v in every chunk
    constexpr int LPR = C::LPR, RPW = 32 / LPR;
    constexpr int RSLOTS = (TM + EPI_WARPS * RPW - 1) / (EPI_WARPS * RPW);
    constexpr int RPF = RSLOTS < 6 ? RSLOTS : 6;                    // readout rows whose Q' operand is requested ahead
    constexpr int NOP = SLOTS > RPF ? SLOTS : RPF;
    const int sub = lane / LPR, hl = lane % LPR;
    int rk[SLOTS], ck[SLOTS];                                       // this thread's units: the same in every chunk / item
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int u = et + k * EPI_THREADS;
      rk[k] = u / UPR;
      ck[k] = 4 * (u - rk[k] * UPR);
    }
    long long* dbg = p.dbg ? p.dbg + (int64_t)blockIdx.x * (MAX_TPC * MAX_LAYERS * 4) : nullptr;
    // h_0 and Q' (outputs of the previous kernels of this forward) are requested by these threads directly, possibly
    // before the producer's first load has landed: every epilogue thread orders itself behind the previous grids
    umma::grid_dep_wait();
    for (int i = 0; i < n_items; ++i) {
      const int j = i % nt, l = i / nt;
      const uint32_t slot = (uint32_t)i & 1u;
      const int tile = tile0 + j;
      const bool readout = l == depth;
      TileAux& ta = aux->t[j];
      const int ecount = ta.info[1], abase = ta.info[2], acount = ta.info[3];
      const float us = __ldg(p.unscale + 1 + l);
      const float skip = (!readout && p.skip[l]) ? __ldg(p.skip[l]) : 1.f;
      __half* o_hi = p.o_hi[(l + 1) & 1];
      __half* o_lo = p.o_lo[(l + 1) & 1];
      float* bias_s = aux->bias_s[slot];
      if (!readout) {
        for (int k = et; k < BN; k += EPI_THREADS) bias_s[k] = n0 + k < H ? __ldg(p.bias[l] + n0 + k) : 0.f;
      } else {
        for (int k = et; k < BN; k += EPI_THREADS) bias_s[k] = n0 + k < H ? __ldg(p.w_ffn + n0 + k) : 0.f;
      }
      // fp32 operand of the epilogue (h0 rows of a bond layer, Q' rows of the readout) for chunk 0: requested before the
      // accumulator is ready, so its L2 latency hides behind the MMAs; later chunks are requested one chunk ahead
      float4 opnd[NOP];
      auto request = [&](int ch, float4 (&dst)[NOP]) {
        const int ncol0 = n0 + ch * CH;
        if (!readout) {
#pragma unroll
          for (int k = 0; k < SLOTS; ++k) {
            const int n = ncol0 + ck[k];
            dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rk[k] < ecount && n < H)
              dst[k] = __ldcg(reinterpret_cast<const float4*>(p.h0 + ((int64_t)tile * TM + rk[k]) * H + n));
          }
        } else {
          const int n = ncol0 + 4 * hl;
#pragma unroll
          for (int k = 0; k < RPF; ++k) {
            const int v = ew * RPW + sub + k * EPI_WARPS * RPW;
            dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < acount && hl < UPR && n < H)
              dst[k] = __ldcg(reinterpret_cast<const float4*>(p.PQ + (int64_t)(abase + v) * (2 * H) + H + n));
          }
        }
      };
      request(0, opnd);
      umma::mbar_wait(umma::smem_u32(&aux->tmem_full[slot]), ((uint32_t)i >> 1) & 1u);
      umma::tc_fence_after_sync();
      if (dbg && et == 0) dbg[i * 4 + 0] = clock64();
      const uint32_t acc = tmem + slot * SLOT_COLS + ((uint32_t)(q * 32) << 16);
      float vmax = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        umma::named_bar_sync(1, EPI_THREADS);                      // the staging rows of the previous chunk are drained
        {
          const int row = q * 32 + lane;
          for (int cc = grp * 8; cc < CH; cc += 8 * (EPI_WARPS / 4)) {
            float v[8];
            umma::tmem_ld_x8(acc + (uint32_t)(ch * CH + cc), v);
            if (C::CAT && !p.fast) {
              float v2[8];
              umma::tmem_ld_x8(acc + (uint32_t)(BN + ch * CH + cc), v2);
              umma::tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] += v2[k];
            } else {
              umma::tmem_ld_wait();
            }
            float4* dst = reinterpret_cast<float4*>(y_s + row * CHP + cc);
            dst[0] = make_float4(v[0] * us, v[1] * us, v[2] * us, v[3] * us);
            dst[1] = make_float4(v[4] * us, v[5] * us, v[6] * us, v[7] * us);
          }
        }
        if (ch == NCH - 1) {                                       // accumulator drained: the MMA warp may reuse the slot
          umma::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(umma::smem_u32(&aux->tmem_empty[slot]));
        }
        float4 cur[NOP];
#pragma unroll
        for (int k = 0; k < NOP; ++k) cur[k] = opnd[k];
        if (NCH > 1 && ch + 1 < NCH) request(ch + 1, opnd);        // next chunk's operand: in flight during this chunk's gather
        umma::named_bar_sync(2, EPI_THREADS);                      // staging rows complete

        const int ncol0 = n0 + ch * CH;                            // first global column of this chunk
        if (!readout) {
#pragma unroll
          for (int k = 0; k < SLOTS; ++k) {
            const int r = rk[k], c = ck[k], n = ncol0 + c;
            if (r < ecount && n < H) {
              const int cnt = ta.cnt_b[r];
              const float4 a4 = cnt <= NBR ? gather_packed<CHP>(y_s + c, ta.nbr_b[r], cnt)
                                           : gather_list<CHP>(y_s + c, ta.idx_l, (int)ta.pb_b[r], (int)ta.full_b[r], r ^ 1);
              const float4 b4 = tcg::ld4(bias_s + ch * CH + c);
              float4 z;
              z.x = a4.x + b4.x + skip * cur[k].x;
              z.y = a4.y + b4.y + skip * cur[k].y;
              z.z = a4.z + b4.z + skip * cur[k].z;
              z.w = a4.w + b4.w + skip * cur[k].w;
              const int64_t orow = (int64_t)tile * TM + r;
              z.x = tcg::act_t<RELU>(z.x, p.act); z.y = tcg::act_t<RELU>(z.y, p.act);
              z.z = tcg::act_t<RELU>(z.z, p.act); z.w = tcg::act_t<RELU>(z.w, p.act);
              vmax = fmaxf(vmax, tcg::amax4(z));
              if (p.fast) {
                const __half2 h01 = __floats2half2_rn(z.x, z.y), h23 = __floats2half2_rn(z.z, z.w);
                uint2 ph;
                ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(o_hi + orow * p.ldo + n) = ph;
              } else {
                tcg::store_split4(z, 1.f, o_hi + orow * p.ldo + n, o_lo + orow * p.ldo + n);
              }
            }
          }
        } else {
          // readout: hv[v] = act(Q'[v] + sum_{k in in(v)} y[k]);  t[v] += hv[v] . w_f over this chunk's columns
          const int c = 4 * hl, n = ncol0 + c;
          const bool lane_on = hl < UPR && n < H;
          const float4 wf4 = lane_on ? tcg::ld4(bias_s + ch * CH + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < RSLOTS; ++k) {
            const int v = ew * RPW + sub + k * EPI_WARPS * RPW;
            if (k * EPI_WARPS * RPW >= acount) break;              // warp-uniform: no row of this step exists
            const bool row_on = v < acount;
            float t = 0.f;
            if (lane_on && row_on) {
              float4 a4 = k < RPF ? cur[k < RPF ? k : 0]
                                  : __ldcg(reinterpret_cast<const float4*>(p.PQ + (int64_t)(abase + v) * (2 * H) + H + n));
              const int cnt = ta.cnt_a[v];
              tcg::add4(a4, cnt <= NBR ? gather_packed<CHP>(y_s + c, ta.nbr_a[v], cnt)
                                       : gather_list<CHP>(y_s + c, ta.idx_l, (int)ta.pb_a[v], cnt, -1));
              t = tcg::act_t<RELU>(a4.x, p.act) * wf4.x;
              t = fmaf(tcg::act_t<RELU>(a4.y, p.act), wf4.y, t);
              t = fmaf(tcg::act_t<RELU>(a4.z, p.act), wf4.z, t);
              t = fmaf(tcg::act_t<RELU>(a4.w, p.act), wf4.w, t);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (hl == 0 && row_on) ta.tat[v] += t;
          }
        }
      }
      if (dbg && et == 0) dbg[i * 4 + 1] = clock64();

      if (!readout) {
        // this CTA's slice of h_{l+1} is stored: make it visible to the peers' TMA loads, then tell every CTA of the cluster
        if (vmax > 60000.f) {                                       // fp16 range of the split: flag the batch and this tile
          atomicOr(p.overflow, 1);
          atomicOr(p.tile_counter + tile, 0x10000);
        }
        __threadfence();
        umma::fence_proxy_async();
        umma::named_bar_sync(3, EPI_THREADS);
        if (et < S) umma::mbar_arrive_remote(umma::smem_u32(&aux->ready[j]), (uint32_t)et);
      } else {
        umma::named_bar_sync(3, EPI_THREADS);                      // every atom's dot product is complete
        const int rx0 = ta.info[4], rxcount = ta.info[5];
        for (int rx = et; rx < rxcount; rx += EPI_THREADS) {
          const int b = rx0 + rx;
          const int v0 = __ldg(p.atom_ptr + b) - abase, v1 = __ldg(p.atom_ptr + b + 1) - abase;
          float s = 0.f;
          for (int v = v0; v < v1; ++v) s += ta.tat[v];            // ascending atom id
          p.partial_out[(int64_t)slice * p.n_rxn + b] = s;
        }
        // the last slice CTA of this tile to arrive adds the slices in a fixed order: deterministic, no extra kernel
        __threadfence();
        umma::named_bar_sync(3, EPI_THREADS);
        if (et == 0) ta.ticket = atomicAdd(p.tile_counter + tile, 1);
        umma::named_bar_sync(3, EPI_THREADS);
        const int ticket = ta.ticket;
        if ((ticket & 0xffff) == S - 1) {
          __threadfence();
          const float bf = __ldg(p.b_ffn);
          // an operand of this tile (or a feature / h_0 of the batch) left the fp16 range: NaN energies, never silent
          const bool poisoned = (ticket & 0x10000) != 0 || (__ldcg(p.overflow) & 3) != 0;
          for (int rx = et; rx < rxcount; rx += EPI_THREADS) {
            const int b = rx0 + rx;
            float s = 0.f;
            for (int k = 0; k < S; ++k) s += __ldcg(p.partial_out + (int64_t)k * p.n_rxn + b);
            p.out[b] = poisoned ? __int_as_float(0x7fc00000) : s + bf;
          }
          if (et == 0) p.tile_counter[tile] = 0;                    // ready for the next forward
        }
      }
      if (dbg && et == 0) dbg[i * 4 + 2] = clock64();
    }
  }

  __syncthreads();
  umma::cluster_sync_all();                       // no CTA leaves while a peer may still touch its barriers
  if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

}  // namespace tcf
