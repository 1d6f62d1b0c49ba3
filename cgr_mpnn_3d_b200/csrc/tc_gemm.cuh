// The fused tcgen05 kernel of the tensor-core engine: TMA-fed FP16x3 GEMM of one 128-row tile with the
// accumulator in TMEM, followed by a shared-memory epilogue (plain / directed-bond gather / readout).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace tcg {

constexpr int TM = 128;                 // rows of a tile (UMMA M)
constexpr int BK = 64;                  // fp16 elements per k-chunk = one 128-byte swizzle row
constexpr int BN = 80;                  // output columns per CTA (UMMA N), multiple of 16
constexpr int BNP = BN + 4;             // padded row of the fp32 staging tiles (conflict-free float4 rows)
constexpr int STAGES = 3;
constexpr int A_BYTES = TM * BK * 2;    // 16 KB
constexpr int B_BYTES = BN * BK * 2;    // 10 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int R_BYTES = TM * BN * 4;    // prefetched fp32 operand of the epilogue (h0 slice / Q slice), dense rows
constexpr int TMEM_COLS = 128;          // power of two >= BN
constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr int AUX_BYTES = 2048;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + R_BYTES + AUX_BYTES + 1024;   // + alignment slack
constexpr int VL = BN / 4;              // lanes that own a float4 column group in the epilogue (20)
static_assert(2 * TM * BNP * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit in the pipeline buffers");
static_assert(BN % 16 == 0 && VL <= 32, "BN must be a multiple of 16 and at most 128");

enum { EPI_PLAIN = 0, EPI_BOND = 1, EPI_READOUT = 2 };

struct TcGemmParams {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  CUtensorMap tmR;              // fp32 [rows, cols] operand prefetched for the epilogue (BOND: h0, READOUT: Q)
  int num_k;                    // k-chunks of BK
  int n_total;                  // real output columns
  int m_rows;                   // real rows (EPI_PLAIN)
  int r_col0;                   // column offset of the R operand inside its tensor (READOUT: H)
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
  int act;
  float dropout_p;
  uint64_t seed;
  uint32_t layer;
  __half* o_hi;                 // next operand, [T*128, ldo]
  __half* o_lo;
  int64_t ldo;
  const float* w_ffn;
  float* partial_out;           // [n_slices, B]
  int64_t n_rxn;
  int* overflow;                // sticky flag: an activation left the fp16 range
};

struct Aux {                    // small per-CTA shared state, lives after the pipeline buffers
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full;
  uint64_t r_full;
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

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = umma::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  float* r_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);          // [TM][BN] dense
  Aux* aux = reinterpret_cast<Aux*>(smem + STAGES * STAGE_BYTES + R_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, slice = blockIdx.y;
  const int n0 = slice * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&aux->full[s]), 1);
      umma::mbar_init(umma::smem_u32(&aux->empty[s]), 1);
    }
    umma::mbar_init(umma::smem_u32(&aux->tmem_full), 1);
    umma::mbar_init(umma::smem_u32(&aux->r_full), 1);
    umma::mbar_fence_init();
    umma::tma_prefetch_desc(&p.tmA_hi);
    umma::tma_prefetch_desc(&p.tmA_lo);
    umma::tma_prefetch_desc(&p.tmB_hi);
    umma::tma_prefetch_desc(&p.tmB_lo);
    if (EPI != EPI_PLAIN) umma::tma_prefetch_desc(&p.tmR);
  }
  if (warp == 1) {
    umma::tmem_alloc(umma::smem_u32(&aux->tmem_base), TMEM_COLS);
    umma::tmem_relinquish();
  }
  if (EPI != EPI_PLAIN && threadIdx.x >= 64 && threadIdx.x < 72)
    aux->info[threadIdx.x - 64] = __ldg(p.tile_info + (int64_t)tile * 8 + (threadIdx.x - 64));
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = aux->tmem_base;

  // ------------------------------------------------------------------ main loop (warp-specialised)
  if (warp == 0) {
    // TMA producer: one elected lane streams A (hi, lo) and B (hi, lo) k-chunks through the ring and
    // prefetches the epilogue's fp32 operand tile
    if (lane == 0 && EPI != EPI_PLAIN) {
      const uint32_t rb = umma::smem_u32(&aux->r_full);
      umma::mbar_arrive_expect_tx(rb, R_BYTES);
      const int row0 = EPI == EPI_BOND ? tile * TM : aux->info[2];
      umma::tma_load_2d(&p.tmR, rb, umma::smem_u32(r_s), p.r_col0 + n0, row0);
    }
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
    // MMA issuer: one lane issues 3 tcgen05.mma per 16-wide k-step (lo.hi + hi.lo + hi.hi)
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

  // per-lane column group of the epilogue: 4 consecutive columns, constant across rows
  const int c = 4 * lane;
  const bool lane_on = lane < VL && (n0 + c) < p.n_total;          // n_total is even; tails handled per element
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), wf4 = bias4;
  if (lane < VL) {
    const int n = n0 + c;
    if (p.bias) {
      bias4.x = n + 0 < p.n_total ? __ldg(p.bias + n + 0) : 0.f;
      bias4.y = n + 1 < p.n_total ? __ldg(p.bias + n + 1) : 0.f;
      bias4.z = n + 2 < p.n_total ? __ldg(p.bias + n + 2) : 0.f;
      bias4.w = n + 3 < p.n_total ? __ldg(p.bias + n + 3) : 0.f;
    }
    if (EPI == EPI_READOUT) {
      wf4.x = n + 0 < p.n_total ? __ldg(p.w_ffn + n + 0) : 0.f;
      wf4.y = n + 1 < p.n_total ? __ldg(p.w_ffn + n + 1) : 0.f;
      wf4.z = n + 2 < p.n_total ? __ldg(p.w_ffn + n + 2) : 0.f;
      wf4.w = n + 3 < p.n_total ? __ldg(p.w_ffn + n + 3) : 0.f;
    }
  }
  const float us = __ldg(p.unscale);
  const float skip = (EPI == EPI_BOND && p.skip) ? __ldg(p.skip) : 1.f;

  // ------------------------------------------------------------------ epilogue (all 8 warps)
  umma::mbar_wait(umma::smem_u32(&aux->tmem_full), 0);
  umma::tc_fence_after_sync();
  float* y_s = reinterpret_cast<float*>(smem);                   // [TM][BNP], aliases the drained pipeline
  float* a_s = y_s + TM * BNP;                                   // [TM][BNP]
  {
    const int q = warp & 3, half = warp >> 2;                    // TMEM lane quarter / column half
    const int row = q * 32 + lane;
    constexpr int COLS_PER_WARP = BN / 2;                        // 40
    float v[COLS_PER_WARP];
#pragma unroll
    for (int cc = 0; cc < COLS_PER_WARP; cc += 8)
      umma::tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * COLS_PER_WARP + cc), v + cc);
    umma::tmem_ld_wait();
#pragma unroll
    for (int cc = 0; cc < COLS_PER_WARP; cc += 4) {
      float4* dst = reinterpret_cast<float4*>(y_s + row * BNP + half * COLS_PER_WARP + cc);
      *dst = make_float4(v[cc] * us, v[cc + 1] * us, v[cc + 2] * us, v[cc + 3] * us);
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();

  if (EPI == EPI_PLAIN) {
    // rows are dense (atoms): out = y + bias, written as coalesced rows
    if (lane < VL) {
      const int n = n0 + c;
      for (int r = warp; r < TM; r += NWARPS) {
        const int64_t row = (int64_t)tile * TM + r;
        if (row >= p.m_rows) break;
        const float4 y = ld4(y_s + r * BNP + c);
        float* o = p.out_f32 + row * p.ldc + n;
        if (n + 3 < p.n_total && (p.ldc & 3) == 0) {
          *reinterpret_cast<float4*>(o) = make_float4(y.x + bias4.x, y.y + bias4.y, y.z + bias4.z, y.w + bias4.w);
        } else {
          if (n + 0 < p.n_total) o[0] = y.x + bias4.x;
          if (n + 1 < p.n_total) o[1] = y.y + bias4.y;
          if (n + 2 < p.n_total) o[2] = y.z + bias4.z;
          if (n + 3 < p.n_total) o[3] = y.w + bias4.w;
        }
      }
    }
  } else if (EPI == EPI_BOND) {
    const int ebase = aux->info[0], ecount = aux->info[1], acount = aux->info[3];
    // a[v] = sum_{k in in(v)} y[k]   (ascending bond id, the reference's accumulation order)
    if (lane < VL) {
      for (int v = warp; v < acount; v += NWARPS) {
        const int pb = aux->ptr_l[v], pe = aux->ptr_l[v + 1];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = pb; q < pe; ++q) {
          const float4 t = ld4(y_s + (int)aux->idx_l[q] * BNP + c);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        *reinterpret_cast<float4*>(a_s + v * BNP + c) = a;
      }
    }
    umma::mbar_wait(umma::smem_u32(&aux->r_full), 0);              // h0 slice has landed (TMA)
    __syncthreads();
    const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    const int H = p.n_total;
    bool ovf = false;
    // z[e] = a[src e] - y[e^1] + b + skip*h0[e];  h' = dropout(act(z));  written as the FP16 (hi, lo) operand
    if (lane_on) {
      const int n = n0 + c;
#pragma unroll 2
      for (int j = warp; j < ecount; j += NWARPS) {
        const int64_t r = (int64_t)tile * TM + j;
        const float4 av = ld4(a_s + (int)aux->src_l[j] * BNP + c);
        const float4 yr = ld4(y_s + (j ^ 1) * BNP + c);
        const float4 h0v = ld4(r_s + j * BN + c);
        float z[4] = {av.x - yr.x + bias4.x + skip * h0v.x, av.y - yr.y + bias4.y + skip * h0v.y,
                      av.z - yr.z + bias4.z + skip * h0v.z, av.w - yr.w + bias4.w + skip * h0v.w};
        __half hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = cgr_act(z[i], p.act);
          if (p.dropout_p > 0.f)
            v = cgr_dropout_keep(p.seed, p.layer, (uint64_t)(ebase + j) * (uint64_t)H + (uint64_t)(n + i), p.dropout_p)
                    ? v * keep_scale : 0.f;
          if (n + i >= H) v = 0.f;
          ovf |= fabsf(v) > 60000.f;
          split_f16(v, hi[i], lo[i]);
        }
        // columns beyond H inside the last 4-group fall in the operand's K padding (never read back: the
        // tensor map's extent is H) -- ldo >= round_up(H, 64)
        uint2 ph, pl;
        ph.x = (uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16);
        ph.y = (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16);
        pl.x = (uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16);
        pl.y = (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16);
        *reinterpret_cast<uint2*>(p.o_hi + r * p.ldo + n) = ph;
        *reinterpret_cast<uint2*>(p.o_lo + r * p.ldo + n) = pl;
      }
    }
    if (ovf) atomicOr(p.overflow, 1);
  } else {
    // readout: hv[v] = act(Q[v] + sum_{k in in(v)} y[k]);  t[v] = hv[v] . w_f (this CTA's columns)
    const int abase = aux->info[2], acount = aux->info[3], rx0 = aux->info[4], rxcount = aux->info[5];
    umma::mbar_wait(umma::smem_u32(&aux->r_full), 0);              // Q slice has landed (TMA)
    for (int v = warp; v < acount; v += NWARPS) {
      float t = 0.f;
      if (lane_on) {
        const int pb = aux->ptr_l[v], pe = aux->ptr_l[v + 1];
        float4 a = ld4(r_s + v * BN + c);
        for (int q = pb; q < pe; ++q) {
          const float4 y = ld4(y_s + (int)aux->idx_l[q] * BNP + c);
          a.x += y.x; a.y += y.y; a.z += y.z; a.w += y.w;
        }
        t = cgr_act(a.x, p.act) * wf4.x;
        t = fmaf(cgr_act(a.y, p.act), wf4.y, t);
        t = fmaf(cgr_act(a.z, p.act), wf4.z, t);
        t = fmaf(cgr_act(a.w, p.act), wf4.w, t);
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

}  // namespace tcg
