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
//   warp 0        TMA producer: weight chunks are requested as soon as a stage is free (they never depend on a peer),
//                 the activation chunks once the tile's previous layer is complete cluster-wide
//   warp 1        tcgen05.mma issuer, accumulators double-buffered in TMEM (2 x 256 columns)
//   warp 2        TMEM allocation
//   warps 4..7    staging: TMEM -> fp32 rows of a two-buffer ring in shared memory, 32 columns per chunk
//   warps 8..23   gather: staged rows -> gather / activation / split -> global; they never touch TMEM
// The staging and gather warps meet through mbarriers only (full / empty per ring buffer): no CTA-wide barrier per
// chunk, a slow gather warp is absorbed by the second buffer, and the accumulator is released as soon as its last
// chunk is staged.  With two tiles per cluster the epilogue of tile X, layer l overlaps the MMAs of tile Y, layer l,
// and so on: the dependency chain of one tile hides behind the other tile's work.
#pragma once
#include "tc_gemm.cuh"

namespace tcf {

using tcg::A_BYTES;
using tcg::BK;
using tcg::NBR;
using tcg::TM;

constexpr int MAX_LAYERS = 16;                 // bond layers + readout
constexpr int STG_WARPS = 4;                   // one per TMEM lane quarter
constexpr int GAT_WARPS = 16;
constexpr int GAT_THREADS = GAT_WARPS * 32;
constexpr int THREADS = 128 + STG_WARPS * 32 + GAT_THREADS;     // 768
constexpr int ZROW = TM;                       // staging row that stays zero: target of unused neighbour slots
constexpr int SLOT_COLS = 256;                 // TMEM columns per accumulator slot
constexpr int MAX_TPC = 2;                     // tiles per cluster (ping-pong)
constexpr int SMEM_LIMIT = 232448;
constexpr int FASTN = 4;                       // neighbours a packed descriptor holds (rows with more walk the CSR list)
// Debug clock64 stamps (tools/fwd_phase_timing.py) are compiled in only with -DCGR_FWD_STAMPS (CGR_FWD_STAMPS=1 at build
// time): the shipped kernel carries no debug stores in its item loop.
#ifdef CGR_FWD_STAMPS
#define FWD_STAMP(cond, slot, w) \
  do { if (p.dbg && (cond)) p.dbg[(int64_t)blockIdx.x * (MAX_TPC * MAX_LAYERS * 8) + (slot) * 8 + (w)] = clock64(); } while (0)
#else
#define FWD_STAMP(cond, slot, w) do { } while (0)
#endif
#ifndef CGR_REG_LOW
#define CGR_REG_LOW 32
#define CGR_REG_HIGH 112
#endif

template <int BN_>
struct FCfg {
  static constexpr int BN = BN_;
  static constexpr bool CAT = 2 * BN <= 256;                   // A_hi x [B_hi ; B_lo] as one MMA of N = 2 BN
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int CH = 32;                                 // columns per staged chunk
  static constexpr int NCH = (BN + CH - 1) / CH;
  static constexpr int UPR = CH / 4;                           // float4 column groups of a chunk: 8 lanes per row
  static constexpr int RPP = GAT_THREADS / UPR;                // rows the gather threads cover per pass: 64
  static constexpr int SLOTS = TM / RPP;                       // passes: units per thread and chunk: 2
  static constexpr int CHP = CH + 4;                           // staging row pitch: 36 floats (conflict-free rows)
  static constexpr int YBUF_BYTES = ((TM + 1) * CHP * 4 + 127) / 128 * 128;     // one ring buffer, + the zero row
  static constexpr int Y_BYTES = 2 * YBUF_BYTES;
  static constexpr int AUX_BYTES = 16384;
  static constexpr int FIT = (SMEM_LIMIT - 1024 - Y_BYTES - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 4 ? 4 : FIT;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + Y_BYTES + AUX_BYTES;
  static_assert(BN % 16 == 0 && BN <= 256 && TM % RPP == 0 && 32 % UPR == 0, "bad slice width");
  static_assert(STAGES >= 2, "pipeline needs two stages");
  static_assert((CAT ? 2 * BN : BN) <= SLOT_COLS, "accumulator exceeds its TMEM slot");
  static_assert((TM + 1) * CHP * 4 < 65536, "staging byte offsets must fit 16 bits");
};

// bytes of the staging ring of a slice width (the fused edge initialisation borrows it)
constexpr int ring_bytes(int bn) { return bn > 128 ? FCfg<208>::Y_BYTES : FCfg<80>::Y_BYTES; }

constexpr int MAX_GROUP = 24;                  // batches one launch can take (cgr_gnn_forward_group)

// what differs between the batches of a group launch: operand buffers, index arrays, outputs
struct FwdBatch {
  CUtensorMap tmA_hi[2], tmA_lo[2];            // activation operand, ping-pong: layer l reads buffer l & 1
  __half* o_hi[2];                             // layer l writes buffer (l + 1) & 1, rows tile * 128 + j (lo: + lo_delta)
  int64_t lo_delta;                            // byte distance from a hi buffer to its lo buffer (same for both pairs)
  float* h0;                                   // [T * 128, H] fp32, tile-packed (skip operand)
  const float* ea;                             // fuse_init: bond features [E, fb] (bond id order)
  const float* PQ;                             // [N, 2H] fp32: P' | Q' (edge-init / readout operands)
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
  int n_tiles;
  int group0;                                  // first tile group (cluster) of this batch in the launch
  // training forward (one operand pair per layer is kept for the backward): train_rows = T * 128 > 0; the pairs lie
  // back to back (hi_0 | lo_0 | hi_1 | ...), tmA_hi[0] spans all of them -- layer l reads rows 2 l train_rows (hi)
  // and (2 l + 1) train_rows (lo) of it and writes o_hi[0] + (l + 1) * 2 * lo_delta; hv_out [N, H] fp32 keeps the
  // readout's activation (mask of its backward)
  int64_t train_rows;
  float* hv_out;
};

struct FwdParams {
  FwdBatch bt[MAX_GROUP];                      // kernel parameters (constant bank): one entry for a plain forward
  int n_batches;
  CUtensorMap tmB_hi[MAX_LAYERS], tmB_lo[MAX_LAYERS];   // prepared weights of bond layer l; [depth] = W_os (readout)
  const float* bias[MAX_LAYERS];               // bond-layer biases
  const float* skip[MAX_LAYERS];               // learnable skip scalars (device) or null (= 1)
  int64_t ldo;
  const float* unscale;                        // [1 + l]: 1 / weight scale of layer l's matrix
  const float* wet;                            // fuse_init: W_e^T [fb, H] fp32 (edge_init.weight[:, Fa:]^T)
  int fb, fuse_init;                           // fuse_init: h0 is computed by this kernel (GNN.py:86), not read
  const float* w_ffn;
  const float* b_ffn;
  int depth, H, num_k, act, tiles_per_cluster;
  int fast;                                    // 1: single-pass fp16 (hi halves only), the "fast" precision mode
  long long* dbg;                              // optional [n_cta][MAX_TPC * MAX_LAYERS][8] clock64 stamps (debug)
  int publish_mode;                            // 0: proxy fence + __threadfence + release arrive; 1: no __threadfence
                                               // (the arrive releases at cluster scope); 2: 1 + proxy fence on .global only;
                                               // 3: 2 + no second proxy fence on the reading side
};

struct TileAux {                // per tile of the group: neighbour descriptors (built once, used by every layer)
  int32_t info[8];
  // bond row j: the in-bonds of src(j) EXCEPT the reverse bond j^1 (GNN.py:141 adds it and subtracts it again); atom row
  // v: the in-bonds of v.  nbd_*: staging-row BYTE offsets (row * CHP * 4) of neighbours 1..4, nb2_*: of neighbours
  // 5..8 (16 bits each), ascending bond id; unused slots point at the zero row, so a gather is four unconditional
  // loads.  Rows with more than 2 * FASTN neighbours walk their CSR list (idx_l).
  uint4 nbd_b[TM];
  uint4 nbd_a[TM];
  uint2 nb2_b[TM];
  uint2 nb2_a[TM];
  uint16_t pb_b[TM];            // CSR offset of the row's full in-bond list
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
  uint64_t st_full[2];          // staging ring: buffer b holds a complete chunk (the STG_WARPS arrived)
  uint64_t st_empty[2];         // every gather warp has consumed buffer b
  uint64_t ready[MAX_TPC];      // tile j's operand of the next layer is complete in every CTA of the cluster
                                // (one arrival per gather warp of every CTA)
  uint32_t tmem_base;
  float us[MAX_LAYERS + 1];     // 1 / weight scale of every layer's matrix (staged once: no global load per item)
  float skipv[MAX_LAYERS];      // learnable skip scalars (1 where the layer has none)
  TileAux t[MAX_TPC];
};
static_assert(sizeof(Aux) <= 16384, "Aux too large");

// four consecutive floats of a parameter vector (16-byte aligned in practice; parameters may also be views)
__device__ __forceinline__ float4 ldg4(const float* p) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// Packed fp32 arithmetic (FADD2 / FFMA2 of sm_100: two IEEE operations per instruction on an aligned register pair --
// the halves of a float4 loaded with one 128-bit access are such pairs).  Results are bit-identical to the scalar
// forms; the gather warps are issue-bound, so halving their adds is worth a fifth of their instructions.
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 sub2(const float2 a, const float2 b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t*>(&a)),
      "l"(*reinterpret_cast<const uint64_t*>(&b)), "l"(*reinterpret_cast<const uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 lo2(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat4(const float2 a, const float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ float4 add4v(const float4 a, const float4 b) { return cat4(add2(lo2(a), lo2(b)), add2(hi2(a), hi2(b))); }

// sum of the four staged rows a descriptor names (byte offsets from y_b): four independent loads, no branches, a
// two-level add tree
__device__ __forceinline__ float4 ldb4(const char* y_b, uint32_t off) { return *reinterpret_cast<const float4*>(y_b + off); }
__device__ __forceinline__ float4 sum4(const float4 v0, const float4 v1, const float4 v2, const float4 v3) {
  return add4v(add4v(v0, v1), add4v(v2, v3));
}
__device__ __forceinline__ float4 gather4(const char* y_b, uint4 o) {
  return sum4(ldb4(y_b, o.x), ldb4(y_b, o.y), ldb4(y_b, o.z), ldb4(y_b, o.w));
}
__device__ __forceinline__ float4 gather4p(const char* y_b, uint2 o) {     // packed: 16 bits per offset
  return sum4(ldb4(y_b, o.x & 0xffffu), ldb4(y_b, o.x >> 16), ldb4(y_b, o.y & 0xffffu), ldb4(y_b, o.y >> 16));
}
// FP16 (hi, lo) split of one float4 column group, stored as two 8-byte words; vmax2 tracks max |hi| of everything this
// thread stored (the fp16-range guard: a value past the range rounds to a hi of inf)
__device__ __forceinline__ void store_split4_track(const float4 z, __half* hi, __half* lo, __half2& vmax2) {
  const __half2 hi01 = __floats2half2_rn(z.x, z.y), hi23 = __floats2half2_rn(z.z, z.w);
  const float2 r01 = sub2(lo2(z), __half22float2(hi01)), r23 = sub2(hi2(z), __half22float2(hi23));
  const __half2 lo01 = __floats2half2_rn(r01.x, r01.y), lo23 = __floats2half2_rn(r23.x, r23.y);
  uint2 ph, pl;
  ph.x = *reinterpret_cast<const uint32_t*>(&hi01); ph.y = *reinterpret_cast<const uint32_t*>(&hi23);
  pl.x = *reinterpret_cast<const uint32_t*>(&lo01); pl.y = *reinterpret_cast<const uint32_t*>(&lo23);
  *reinterpret_cast<uint2*>(hi) = ph;
  *reinterpret_cast<uint2*>(lo) = pl;
  vmax2 = __hmax2(vmax2, __hmax2(__habs2(hi01), __habs2(hi23)));
}
// rows with more than 2 * FASTN neighbours: walk the CSR list, skipping row `skip`
template <int CHP>
__device__ __noinline__ float4 gather_list(const char* y_b, const uint8_t* idx_l, int pb, int n, int skip) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < n; ++t) {
    const int k = idx_l[pb + t];
    if (k != skip) a = add4v(a, ldb4(y_b, (uint32_t)(k * CHP * 4)));
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
  // the launch's clusters are the tile groups of its batches, batch after batch
  int group = (int)blockIdx.x / S, bi = 0;
  while (bi + 1 < p.n_batches && group >= p.bt[bi + 1].group0) ++bi;
  const FwdBatch& B = p.bt[bi];
  group -= B.group0;
  const int n0 = slice * BN;
  const int H = p.H, depth = p.depth;
  int n_eff = H - n0;                                              // columns this slice owns, rounded to the MMA granularity
  n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
  const int tile0 = group * p.tiles_per_cluster;
  int nt = B.n_tiles - tile0;
  nt = nt < p.tiles_per_cluster ? nt : p.tiles_per_cluster;
  const int n_items = nt * (depth + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->tmem_full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->tmem_empty[s]), STG_WARPS);
      umma::mbar_init(umma::smem_u32(&aux->st_full[s]), STG_WARPS);
      umma::mbar_init(umma::smem_u32(&aux->st_empty[s]), GAT_WARPS);
    }
    for (int j = 0; j < MAX_TPC; ++j) umma::mbar_init(umma::smem_u32(&aux->ready[j]), (uint32_t)(S * GAT_WARPS));
    umma::mbar_fence_init();
    umma::tma_prefetch_desc(&B.tmA_hi[0]);
    umma::tma_prefetch_desc(&B.tmA_lo[0]);
    umma::tma_prefetch_desc(&p.tmB_hi[0]);
    umma::tma_prefetch_desc(&p.tmB_lo[0]);
  }
  if (warp == 2) {
    umma::tmem_alloc(umma::smem_u32(&aux->tmem_base), 512);
    umma::tmem_relinquish();
  }
  if (threadIdx.x >= 96 && threadIdx.x < 96 + 8 * MAX_TPC) {
    const int j = (threadIdx.x - 96) >> 3, k = (threadIdx.x - 96) & 7;
    aux->t[j].info[k] = j < nt ? __ldg(B.tile_info + (int64_t)(tile0 + j) * 8 + k) : 0;
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
    for (int i = threadIdx.x; i < ecount; i += THREADS) ta.idx_l[i] = (uint8_t)(__ldg(B.in_idx + ebase + i) - ebase);
    for (int rr = threadIdx.x; rr < 2 * TM; rr += THREADS) {       // every row gets a descriptor: dead rows read zeros
      const bool bond = rr < TM;
      const int r = bond ? rr : rr - TM;
      const bool live = r < (bond ? ecount : acount);
      const int a = !live ? 0 : (bond ? __ldg(B.src + ebase + r) : abase + r);
      const int skip = bond ? (r ^ 1) : -1;
      const int pb = live ? __ldg(B.in_ptr + a) : ebase, pe = live ? __ldg(B.in_ptr + a + 1) : ebase;
      uint32_t o[2 * FASTN];
#pragma unroll
      for (int t = 0; t < 2 * FASTN; ++t) o[t] = ZROW * CHP * 4;
      int cnt = 0;
      for (int t = pb; t < pe; ++t) {
        const int k = __ldg(B.in_idx + t) - ebase;
        if (k == skip) continue;
#pragma unroll
        for (int u = 0; u < 2 * FASTN; ++u)
          if (cnt == u) o[u] = (uint32_t)(k * CHP * 4);
        ++cnt;
      }
      const uint4 d1 = make_uint4(o[0], o[1], o[2], o[3]);
      const uint2 d2 = make_uint2(o[4] | (o[5] << 16), o[6] | (o[7] << 16));
      if (bond) {
        ta.nbd_b[r] = d1; ta.nb2_b[r] = d2; ta.pb_b[r] = (uint16_t)(pb - ebase);
        ta.cnt_b[r] = (uint8_t)(cnt > 255 ? 255 : cnt); ta.full_b[r] = (uint8_t)(pe - pb > 255 ? 255 : pe - pb);
      } else {
        ta.nbd_a[r] = d1; ta.nb2_a[r] = d2; ta.pb_a[r] = (uint16_t)(pb - ebase);
        ta.cnt_a[r] = (uint8_t)(cnt > 255 ? 255 : cnt);
      }
    }
    for (int v = threadIdx.x; v < TM; v += THREADS) ta.tat[v] = 0.f;
  }
  for (int k = threadIdx.x; k < 2 * CHP; k += THREADS)              // the zero row of both ring buffers
    y_s[(k / CHP) * (C::YBUF_BYTES / 4) + ZROW * CHP + (k % CHP)] = 0.f;
  // per-layer scalars staged once: 1 / weight scale (written by the weight preparation, complete before the first
  // kernel of this forward started), learnable skip
  if ((int)threadIdx.x <= depth) aux->us[threadIdx.x] = __ldg(p.unscale + 1 + threadIdx.x);
  if ((int)threadIdx.x < depth) aux->skipv[threadIdx.x] = p.skip[threadIdx.x] ? __ldg(p.skip[threadIdx.x]) : 1.f;
  __syncthreads();

  // 640 threads x 96 registers are allocated at launch: warps 0..3 keep 32, the 16 epilogue warps grow to 112
#ifdef CGR_SETMAXNREG   // measured: ptxas spills ~1.3 KB per thread with the split budget; off
  if (warp < 4) umma::reg_dec<CGR_REG_LOW>(); else umma::reg_inc<CGR_REG_HIGH>();
#endif

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
          if (kc == 0 && (l > 0 || p.fuse_init)) {
            // the S slices of this tile's previous layer have been stored (every CTA of the cluster arrived)
            umma::mbar_wait_cluster(umma::smem_u32(&aux->ready[j]), (uint32_t)(l - 1 + p.fuse_init) & 1u);
            if (p.publish_mode < 3) umma::fence_proxy_async();
            FWD_STAMP(true, i, 3);
          }
          if (B.train_rows) {
            const int row = tile * TM + (int)(2 * l * B.train_rows);
            umma::tma_load_2d(&B.tmA_hi[0], full, st, kc * BK, row);
            if (!p.fast) umma::tma_load_2d(&B.tmA_hi[0], full, st + A_BYTES, kc * BK, row + (int)B.train_rows);
          } else {
            umma::tma_load_2d(&B.tmA_hi[buf], full, st, kc * BK, tile * TM);
            if (!p.fast) umma::tma_load_2d(&B.tmA_lo[buf], full, st + A_BYTES, kc * BK, tile * TM);
          }
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
        FWD_STAMP(true, i, 4);
        const uint32_t acc = tmem + slot * SLOT_COLS;
        for (int kc = 0; kc < p.num_k; ++kc, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          umma::mbar_wait(umma::smem_u32(&aux->full[s]), ph);
          umma::tc_fence_after_sync();
          FWD_STAMP(kc == 0, i, 7);
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
  } else if (warp >= 4 && warp < 4 + STG_WARPS) {
    // ------------------------------------------------------------------ staging warps: TMEM -> ring buffer
    const int q = warp & 3;                                         // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    uint32_t c = 0;                                                 // chunks staged so far (ring position)
    for (int i = 0; i < n_items; ++i) {
      const uint32_t slot = (uint32_t)i & 1u;
      const float us = aux->us[nt == 2 ? (i >> 1) : i];
      umma::mbar_wait(umma::smem_u32(&aux->tmem_full[slot]), ((uint32_t)i >> 1) & 1u);
      umma::tc_fence_after_sync();
      FWD_STAMP(warp == 4 && lane == 0, i, 5);
      const uint32_t acc = tmem + slot * SLOT_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch, ++c) {
        const int cw = (ch + 1) * CH <= BN ? CH : BN - ch * CH;    // columns of this chunk (the last one may be shorter)
        const uint32_t buf = c & 1u;
        // use u of a buffer waits for the gather warps' release of use u - 1; with the fused edge initialisation the
        // ring's first release is the end of that phase (it borrows the ring), so the phases shift by one
        umma::mbar_wait(umma::smem_u32(&aux->st_empty[buf]), ((c >> 1) & 1u) ^ 1u ^ (uint32_t)p.fuse_init);
        float* dst_row = y_s + buf * (C::YBUF_BYTES / 4) + row * CHP;
        if (C::CAT && !p.fast) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {                             // two 8-column groups at a time: hi.hi + hi.lo parts
            float v[2][8], v2[2][8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int cc = (2 * h + g) * 8;
              if (cc < cw) {
                umma::tmem_ld_x8(acc + (uint32_t)(ch * CH + cc), v[g]);
                umma::tmem_ld_x8(acc + (uint32_t)(BN + ch * CH + cc), v2[g]);
              }
            }
            umma::tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int cc = (2 * h + g) * 8;
              if (cc < cw) {
                float4* dst = reinterpret_cast<float4*>(dst_row + cc);
                dst[0] = make_float4((v[g][0] + v2[g][0]) * us, (v[g][1] + v2[g][1]) * us, (v[g][2] + v2[g][2]) * us,
                                     (v[g][3] + v2[g][3]) * us);
                dst[1] = make_float4((v[g][4] + v2[g][4]) * us, (v[g][5] + v2[g][5]) * us, (v[g][6] + v2[g][6]) * us,
                                     (v[g][7] + v2[g][7]) * us);
              }
            }
          }
        } else {
          float v[CH / 8][8];
#pragma unroll
          for (int g = 0; g < CH / 8; ++g)                           // all loads of the chunk in flight, one wait
            if (g * 8 < cw) umma::tmem_ld_x8(acc + (uint32_t)(ch * CH + g * 8), v[g]);
          umma::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < CH / 8; ++g) {
            if (g * 8 < cw) {
              float4* dst = reinterpret_cast<float4*>(dst_row + g * 8);
              dst[0] = make_float4(v[g][0] * us, v[g][1] * us, v[g][2] * us, v[g][3] * us);
              dst[1] = make_float4(v[g][4] * us, v[g][5] * us, v[g][6] * us, v[g][7] * us);
            }
          }
        }
        if (ch == NCH - 1) umma::tc_fence_before_sync();            // accumulator drained: the MMA warp may reuse it
        __syncwarp();
        if (lane == 0) {
          if (ch == NCH - 1) umma::mbar_arrive(umma::smem_u32(&aux->tmem_empty[slot]));
          umma::mbar_arrive(umma::smem_u32(&aux->st_full[buf]));    // release: the warp's rows are in the buffer
        }
      }
      FWD_STAMP(warp == 4 && lane == 0, i, 6);
    }
  } else if (warp >= 4 + STG_WARPS) {
    // ------------------------------------------------------------------ gather warps
    constexpr int RPP = C::RPP, SLOTS = C::SLOTS;
    const int et = (int)threadIdx.x - (128 + STG_WARPS * 32);       // 0 .. GAT_THREADS - 1
    // thread et owns column group cg (4 columns) of the rows r0, r0 + RPP in every chunk: 8 lanes per row, so a
    // warp reads four 128-byte row segments per load (no bank conflicts) and writes 64-byte segments
    const int cg = et & (UPR - 1), r0 = et / UPR;
    const uint32_t col_b = (uint32_t)(16 * cg);                    // this thread's column group, bytes into a staging row
    const uint32_t h0_row_step = (uint32_t)(RPP * H) * 4u;         // bytes between this thread's two h0 rows
    const uint32_t pq_row_step = (uint32_t)(RPP * 2 * H) * 4u;     // ... its two Q' rows
    const uint32_t o_row_step = (uint32_t)RPP * (uint32_t)p.ldo * 2u;   // ... its two (hi, lo) output rows
    auto stamp = [&](int i, int w) { FWD_STAMP(et == 0, i, w); (void)i; (void)w; };   // debug: clock64 phase stamps
    // item i of this CTA = (tile j of the group, layer l); nt is 1 or 2
    auto item_tile = [&](int i) { return nt == 2 ? (i & 1) : 0; };
    auto item_layer = [&](int i) { return nt == 2 ? (i >> 1) : i; };
    // h_0 and Q' (outputs of the previous kernels of this forward) are read by these threads directly
    umma::grid_dep_wait();

    // operands of one chunk, requested one chunk ahead (their L2 latency never shows): the fp32 rows added after the
    // gather (h0 of a bond layer, Q' of the readout) and the per-column vector (bias / w_ffn)
    float4 opnd[SLOTS], c4;
    auto request = [&](int i, int ch) {
      const int j = item_tile(i), l = item_layer(i);
      const TileAux& ta = aux->t[j];
      const int n = n0 + ch * CH + 4 * cg;
      const bool col_on = ch * CH + 4 * cg < BN && n < H;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) opnd[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      c4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!col_on) return;
      if (l < depth) {
        c4 = ldg4(p.bias[l] + n);
        const int ecount = ta.info[1];
        const char* hp = reinterpret_cast<const char*>(B.h0 + ((int64_t)(tile0 + j) * TM + r0) * H + n);
#pragma unroll
        for (int k = 0; k < SLOTS; ++k)
          if (r0 + k * RPP < ecount) opnd[k] = __ldcg(reinterpret_cast<const float4*>(hp + (uint32_t)k * h0_row_step));
      } else {
        c4 = ldg4(p.w_ffn + n);
        const int abase = ta.info[2], acount = ta.info[3];
        const char* qp = reinterpret_cast<const char*>(B.PQ + (int64_t)(abase + r0) * (2 * H) + H + n);
#pragma unroll
        for (int k = 0; k < SLOTS; ++k)
          if (r0 + k * RPP < acount) opnd[k] = __ldcg(reinterpret_cast<const float4*>(qp + (uint32_t)k * pq_row_step));
      }
    };
    stamp(MAX_TPC * MAX_LAYERS - 1, 0);
    if (p.fuse_init) {
      // ---- edge initialisation for this slice's columns of the group's tiles (GNN.py:86):
      //      h0[e] = act(P'[src e] + ea[e] . W_e^T), written as the fp32 skip operand and the (hi, lo) operand of layer 0.
      // The ring memory is free until the first accumulator is staged: it holds W_e^T's slice and the tiles' bond
      // features meanwhile.  8 lanes per row; a lane requests all its column groups of the row before it computes.
      const int fb = p.fb;
      // source atoms of this thread's rows (-1: dead row), requested first: the latency hides behind the staging below
      int srow[MAX_TPC * SLOTS];
#pragma unroll
      for (int q = 0; q < MAX_TPC * SLOTS; ++q) {
        const int jj = q / SLOTS, r = r0 + (q % SLOTS) * RPP;
        srow[q] = (jj < nt && r < aux->t[jj].info[1]) ? __ldg(B.src + aux->t[jj].info[0] + r) : -1;
      }
      float* wet_s = y_s;                                           // [fb][BN]
      float* ea_s = y_s + fb * BN;                                  // [nt * TM][fb]
      for (int t = et; t < fb * (BN / 4); t += GAT_THREADS) {
        const int k = t / (BN / 4), c = 4 * (t % (BN / 4));
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + c < H) w = ldg4(p.wet + (size_t)k * H + n0 + c);
        *reinterpret_cast<float4*>(wet_s + k * BN + c) = w;
      }
      {
        // 16 threads per bond row; every thread's loads are issued before its first store (memory-level parallelism)
        constexpr int NP = (MAX_TPC * TM) / (GAT_THREADS / 16);      // 8 passes cover two tiles
        float ev[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const int jr = (et >> 4) + q * (GAT_THREADS / 16), k = et & 15;
          ev[q] = 0.f;
          if (jr < nt * TM && k < fb) {
            const TileAux& ta = aux->t[jr / TM];
            const int r = jr % TM;
            if (r < ta.info[1]) ev[q] = __ldg(B.ea + (size_t)(ta.info[0] + r) * fb + k);
          }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const int jr = (et >> 4) + q * (GAT_THREADS / 16), k = et & 15;
          if (jr < nt * TM && k < fb) ea_s[jr * fb + k] = ev[q];
        }
        for (int jr = et >> 4; fb > 16 && jr < nt * TM; jr += GAT_THREADS / 16) {      // wider bond features: the rest
          const TileAux& ta = aux->t[jr / TM];
          const int r = jr % TM;
          for (int k = 16 + (et & 15); k < fb; k += 16)
            ea_s[jr * fb + k] = r < ta.info[1] ? __ldg(B.ea + (size_t)(ta.info[0] + r) * fb + k) : 0.f;
        }
      }
      umma::named_bar_sync(3, GAT_THREADS);
      stamp(MAX_TPC * MAX_LAYERS - 1, 2);
      float imax = 0.f;
      constexpr int NI = (BN / 4 + UPR - 1) / UPR;                  // column groups of the slice per lane
      // One step = (tile, column group i of this lane) for the thread's SLOTS rows.  The P' operands of step q + 1 are
      // requested before step q is computed, so the L2 latency of the gathered rows never shows (it used to be paid
      // once per row: 45 k cycles for two tiles).
      auto fetch = [&](int q, float4 (&dst)[SLOTS]) {
        const int jj = q / NI, i = q - jj * NI;
        const int c = 4 * (cg + UPR * i);
        const bool on = c < BN && n0 + c < H;
#pragma unroll
        for (int u = 0; u < SLOTS; ++u) {
          const int sa = jj ? srow[SLOTS + u] : srow[u];
          dst[u] = (on && sa >= 0) ? __ldcg(reinterpret_cast<const float4*>(B.PQ + (size_t)sa * (2 * H) + n0 + c))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 cur[SLOTS], nxt[SLOTS];
      const int nq = nt * NI;
      fetch(0, cur);
#pragma unroll 1
      for (int q = 0; q < nq; ++q) {
        if (q + 1 < nq) fetch(q + 1, nxt);
        const int jj = q / NI, i = q - jj * NI;
        const int c = 4 * (cg + UPR * i);
        if (c < BN && n0 + c < H) {
          const float* wr = wet_s + c;
          const float* e0 = ea_s + (jj * TM + r0) * fb;
#pragma unroll 2
          for (int k = 0; k < fb; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(wr + k * BN);
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const float x = e0[u * RPP * fb + k];                  // dead rows hold zeros
              cur[u].x = fmaf(x, w.x, cur[u].x); cur[u].y = fmaf(x, w.y, cur[u].y);
              cur[u].z = fmaf(x, w.z, cur[u].z); cur[u].w = fmaf(x, w.w, cur[u].w);
            }
          }
#pragma unroll
          for (int u = 0; u < SLOTS; ++u) {
            if ((jj ? srow[SLOTS + u] : srow[u]) >= 0) {
              float4 v = cur[u];
              v.x = tcg::act_t<RELU>(v.x, p.act); v.y = tcg::act_t<RELU>(v.y, p.act);
              v.z = tcg::act_t<RELU>(v.z, p.act); v.w = tcg::act_t<RELU>(v.w, p.act);
              imax = fmaxf(imax, tcg::amax4(v));
              const size_t row = (size_t)(tile0 + jj) * TM + r0 + u * RPP;
              *reinterpret_cast<float4*>(B.h0 + row * H + n0 + c) = v;
              char* od = reinterpret_cast<char*>(B.o_hi[0]) + (row * (size_t)p.ldo + n0 + c) * 2;
              tcg::store_split4(v, 1.f, reinterpret_cast<__half*>(od), reinterpret_cast<__half*>(od + B.lo_delta));
            }
          }
        }
        if (i == NI - 1) {
          stamp(MAX_TPC * MAX_LAYERS - 1, 3 + 2 * jj);
          // publish this tile's slice of the layer-0 operand (same protocol as a layer's output, below)
          if (p.publish_mode < 2) umma::fence_proxy_async(); else umma::fence_proxy_async_global();
          __syncwarp();
          if (lane < S) {
            if (p.publish_mode == 0) __threadfence();
            umma::mbar_arrive_remote(umma::smem_u32(&aux->ready[jj]), (uint32_t)lane);
          }
          stamp(MAX_TPC * MAX_LAYERS - 1, 4 + 2 * jj);
        }
#pragma unroll
        for (int u = 0; u < SLOTS; ++u) cur[u] = nxt[u];
      }
      if (imax > 60000.f) atomicOr(B.overflow, 1);                  // fp16 range of the split (as edge_init flags it)
      // the ring goes back to its own use: its zero rows again, and this CTA's h0 columns are visible to every
      // gather thread (they are read back as the skip operand)
      umma::named_bar_sync(3, GAT_THREADS);
      for (int k = et; k < 2 * CHP; k += GAT_THREADS)
        y_s[(k / CHP) * (C::YBUF_BYTES / 4) + ZROW * CHP + (k % CHP)] = 0.f;
      __threadfence_block();
      umma::named_bar_sync(3, GAT_THREADS);
      if (lane == 0) {                                              // the staging warps may use the ring from here on
        umma::mbar_arrive(umma::smem_u32(&aux->st_empty[0]));
        umma::mbar_arrive(umma::smem_u32(&aux->st_empty[1]));
      }
    }
    stamp(MAX_TPC * MAX_LAYERS - 1, 1);
    request(0, 0);

    uint32_t c = 0;                                                 // chunks consumed so far (ring position)
    for (int i = 0; i < n_items; ++i) {
      const int j = item_tile(i), l = item_layer(i);
      const int tile = tile0 + j;
      const bool readout = l == depth;
      TileAux& ta = aux->t[j];
      const int abase = ta.info[2];
      const int count = readout ? ta.info[3] : ta.info[1];          // live rows of this item: atoms / bonds
      const float skip = readout ? 1.f : aux->skipv[l];
      const uint4* nbd = readout ? ta.nbd_a : ta.nbd_b;
      const uint2* nb2 = readout ? ta.nb2_a : ta.nb2_b;
      // this thread's unit (row r0, column group cg) of the output operand; the lo rows sit B.lo_delta bytes further
      char* oh_item = (B.train_rows ? reinterpret_cast<char*>(B.o_hi[0]) + (int64_t)(l + 1) * 2 * B.lo_delta
                                    : reinterpret_cast<char*>(B.o_hi[(l + 1) & 1])) +
                      (((size_t)tile * TM + r0) * (size_t)p.ldo + n0 + 4 * cg) * 2;
      asm volatile("" : "+l"(oh_item));                            // opaque: kept in registers, not recomputed per unit
      // this thread's rows: which exist, which have more than FASTN / 2 FASTN neighbours
      uint32_t valid = 0, slow = 0, vslow = 0;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) {
        const int r = r0 + k * RPP;
        if (r < count) {
          valid |= 1u << k;
          const int cn = readout ? ta.cnt_a[r] : ta.cnt_b[r];
          if (cn > FASTN) slow |= 1u << k;
          if (cn > 2 * FASTN) vslow |= 1u << k;
        }
      }
      __half2 vmax2 = __float2half2_rn(0.f);                        // max |hi| of what this thread stores (fp16-range guard)
      float tsum[SLOTS];                                            // readout: this thread's part of hv[v] . w_f
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) tsum[k] = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch, ++c) {
        const int n = n0 + ch * CH + 4 * cg;
        const bool col_on = ch * CH + 4 * cg < BN && n < H;
        const uint32_t buf = c & 1u;
        umma::mbar_wait(umma::smem_u32(&aux->st_full[buf]), (c >> 1) & 1u);
        if (ch == 0) stamp(i, 0);
        const char* y_bc = reinterpret_cast<const char*>(y_s) + buf * C::YBUF_BYTES + col_b;
        if (col_on) {
          char* oh = oh_item + ch * (CH * 2);
#pragma unroll
          for (int k = 0; k < SLOTS; ++k) {
            const int r = r0 + k * RPP;
            float4 a4 = gather4(y_bc, nbd[r]);                      // dead rows and unused slots read the zero row
            if (slow & (1u << k)) a4 = add4v(a4, gather4p(y_bc, nb2[r]));         // neighbours 5..8 (one row in 25)
            if (vslow & (1u << k))                                  // an atom with ten or more bonds: walk the CSR list
              a4 = readout ? gather_list<CHP>(y_bc, ta.idx_l, (int)ta.pb_a[r], (int)ta.cnt_a[r], -1)
                           : gather_list<CHP>(y_bc, ta.idx_l, (int)ta.pb_b[r], (int)ta.full_b[r], r ^ 1);
            if (!readout) {
              // z[e] = sum_{k in in(src e), k != e^1} y[k] + b + skip * h0[e];  h' = act(z) -> next operand (hi, lo)
              const float2 sk2 = make_float2(skip, skip);
              const float2 p01 = add2(lo2(a4), fma2(sk2, lo2(opnd[k]), lo2(c4)));
              const float2 p23 = add2(hi2(a4), fma2(sk2, hi2(opnd[k]), hi2(c4)));
              float4 z;
              z.x = tcg::act_t<RELU>(p01.x, p.act); z.y = tcg::act_t<RELU>(p01.y, p.act);
              z.z = tcg::act_t<RELU>(p23.x, p.act); z.w = tcg::act_t<RELU>(p23.y, p.act);
              if (valid & (1u << k)) {
                char* od = oh + (uint32_t)k * o_row_step;
                store_split4_track(z, reinterpret_cast<__half*>(od), reinterpret_cast<__half*>(od + B.lo_delta), vmax2);
              }
            } else {
              // readout: hv[v] = act(Q'[v] + sum_{k in in(v)} y[k]);  t[v] += hv[v] . w_f over this thread's columns
              const float4 q4 = add4v(a4, opnd[k]);
              float4 hv;
              hv.x = tcg::act_t<RELU>(q4.x, p.act); hv.y = tcg::act_t<RELU>(q4.y, p.act);
              hv.z = tcg::act_t<RELU>(q4.z, p.act); hv.w = tcg::act_t<RELU>(q4.w, p.act);
              float t = hv.x * c4.x;
              t = fmaf(hv.y, c4.y, t);
              t = fmaf(hv.z, c4.z, t);
              t = fmaf(hv.w, c4.w, t);
              tsum[k] += t;
              if (B.hv_out && (valid & (1u << k)))                   // training: the readout backward's mask
                *reinterpret_cast<float4*>(B.hv_out + (size_t)(abase + r) * H + n) = hv;
            }
          }
        }
        __syncwarp();                                               // every lane's reads of the buffer are done
        if (lane == 0) umma::mbar_arrive(umma::smem_u32(&aux->st_empty[buf]));
        // the operands of the next chunk (or of the next item's first chunk) go in flight now
        if (ch + 1 < NCH) request(i, ch + 1);
        else if (i + 1 < n_items) request(i + 1, 0);
      }
      stamp(i, 1);

      if (!readout) {
        // this CTA's slice of h_{l+1} is stored: make it visible to the peers' TMA loads, then tell every CTA of the cluster
        const float vmax = fmaxf(__low2float(vmax2), __high2float(vmax2));
        if (vmax > 60000.f) {                                       // fp16 range of the split: flag the batch and this tile
          atomicOr(B.overflow, 1);
          atomicOr(B.tile_counter + tile, 0x10000);
        }
        // every gather warp publishes its own rows: the lanes order their stores against the peers' TMA reads (proxy
        // fence) and meet at the warp barrier, then one lane per peer CTA arrives on that CTA's barrier with release
        // semantics at cluster scope (cumulative over what the warp barrier ordered before it; a separate
        // __threadfence() there cost a sequentially consistent fence + an L1 invalidate per item for nothing:
        // publish_mode 0) -- S * GAT_WARPS arrivals complete a layer.
        // No CTA-wide barrier: a warp that is done moves on to the next item's chunks
        if (p.publish_mode < 2) umma::fence_proxy_async(); else umma::fence_proxy_async_global();
        __syncwarp();
        if (lane < S) {
          if (p.publish_mode == 0) __threadfence();
          umma::mbar_arrive_remote(umma::smem_u32(&aux->ready[j]), (uint32_t)lane);
        }
      } else {
        // the 8 lanes of a row add their parts in a fixed tree order: the atom's dot product with w_f
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
          float t = tsum[k];
#pragma unroll
          for (int o = UPR / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (cg == 0) ta.tat[r0 + k * RPP] = t;
        }
        umma::named_bar_sync(3, GAT_THREADS);                      // every atom's dot product is complete
        const int rx0 = ta.info[4], rxcount = ta.info[5];
        for (int rx = et; rx < rxcount; rx += GAT_THREADS) {
          const int b = rx0 + rx;
          const int v0 = __ldg(B.atom_ptr + b) - abase, v1 = __ldg(B.atom_ptr + b + 1) - abase;
          float s = 0.f;
          for (int v = v0; v < v1; ++v) s += ta.tat[v];            // ascending atom id
          B.partial_out[(int64_t)slice * B.n_rxn + b] = s;
        }
        // the last slice CTA of this tile to arrive adds the slices in a fixed order: deterministic, no extra kernel
        __threadfence();
        umma::named_bar_sync(3, GAT_THREADS);
        if (et == 0) ta.ticket = atomicAdd(B.tile_counter + tile, 1);
        umma::named_bar_sync(3, GAT_THREADS);
        const int ticket = ta.ticket;
        if ((ticket & 0xffff) == S - 1) {
          __threadfence();
          const float bf = __ldg(p.b_ffn);
          // an operand of this tile (or a feature / h_0 of the batch) left the fp16 range: NaN energies, never silent
          const bool poisoned = (ticket & 0x10000) != 0 || (__ldcg(B.overflow) & 3) != 0;
          for (int rx = et; rx < rxcount; rx += GAT_THREADS) {
            const int b = rx0 + rx;
            float s = 0.f;
            for (int k = 0; k < S; ++k) s += __ldcg(B.partial_out + (int64_t)k * B.n_rxn + b);
            B.out[b] = poisoned ? __int_as_float(0x7fc00000) : s + bf;
          }
          if (et == 0) B.tile_counter[tile] = 0;                    // ready for the next forward
        }
      }
      stamp(i, 2);
    }
  }

  __syncthreads();
  umma::cluster_sync_all();                       // no CTA leaves while a peer may still touch its barriers
  if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

}  // namespace tcf
