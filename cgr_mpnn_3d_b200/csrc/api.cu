// C ABI of libcgr_b200: stage-level entry points and whole-network forward / backward
// (reference cgr_mpnn_3D/models/GNN.py:76-145; backward = explicit mirror of autograd).
#include <stdarg.h>
#include <string.h>

#include "../../include/cgr_b200.h"
#include "common.cuh"
#include "simt.cuh"
#include "tc.cuh"

static thread_local char g_err[512] = "";

void cgr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch counter and event-based per-stage timing -------------------------------------------
#include <atomic>
#include <mutex>
#include <vector>
#include <string>
namespace {
std::atomic<long long> g_launches{0};
std::atomic<int> g_profile{0};
std::mutex g_prof_mu;
struct ProfRec { std::string name; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof;
}  // namespace

void cgr_note_launch(const char*, cudaStream_t, int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

CgrRange::CgrRange(const char* name_, cudaStream_t st_) : name(name_), st(st_), active(g_profile.load() != 0) {
  if (!active) return;
  ProfRec r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) { active = false; return; }
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
CgrRange::~CgrRange() {
  if (!active) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (size_t i = g_prof.size(); i-- > 0;)
    if (g_prof[i].name == name) { cudaEventRecord(g_prof[i].b, st); break; }
}

extern "C" long long cgr_launch_count(void) { return g_launches.load(); }
extern "C" int cgr_profile_enable(int enable) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_profile.store(enable ? 1 : 0);
  return CGR_OK;
}
extern "C" int cgr_profile_count(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  return (int)g_prof.size();
}
extern "C" int cgr_profile_get(int i, char* name_out, int name_cap, float* ms_out) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  CGR_CHECK_ARG(i >= 0 && i < (int)g_prof.size() && name_out && name_cap > 0 && ms_out, "cgr_profile_get: bad index");
  CGR_CUDA(cudaEventSynchronize(g_prof[i].b));
  CGR_CUDA(cudaEventElapsedTime(ms_out, g_prof[i].a, g_prof[i].b));
  snprintf(name_out, name_cap, "%s", g_prof[i].name.c_str());
  return CGR_OK;
}

extern "C" int cgr_version(void) { return CGR_B200_VERSION; }
extern "C" const char* cgr_last_error_string(void) { return g_err; }

namespace {

struct Carver {   // bump allocator over a caller-provided workspace, 256-byte aligned pieces
  char* base;
  size_t cap, off = 0;
  bool ok = true;
  Carver(void* p, size_t n) : base((char*)p), cap(n) {}
  float* floats(size_t n) {
    size_t bytes = cgr_align_up(n * sizeof(float), 256);
    if (off + bytes > cap) { ok = false; return nullptr; }
    float* r = (float*)(base + off);
    off += bytes;
    return r;
  }
};
size_t fbytes(size_t n) { return cgr_align_up(n * sizeof(float), 256); }

int check_params(const cgr_params_t* p) {
  CGR_CHECK_ARG(p, "null params");
  CGR_CHECK_ARG(p->hidden > 0 && p->depth > 0 && p->fa > 0 && p->fb >= 0, "bad architecture sizes");
  CGR_CHECK_ARG(p->act >= 0 && p->act <= 2, "unknown activation id %d", p->act);
  CGR_CHECK_ARG(p->w_init && p->b_init && p->w_conv && p->b_conv && p->w_e2n && p->b_e2n && p->w_ffn && p->b_ffn,
                "null parameter pointer");
  CGR_CHECK_ARG(!p->use_skip || p->skip, "use_skip set but skip pointers missing");
  return CGR_OK;
}

int check_graph(const cgr_graph_t* g) {
  CGR_CHECK_ARG(g, "null graph");
  CGR_CHECK_ARG(g->n_atoms > 0 && g->n_bonds > 0 && g->n_rxn > 0, "empty batch");
  CGR_CHECK_ARG((g->n_bonds & 1) == 0, "odd number of directed bonds (reference GNN.py:136-138 needs pairs)");
  CGR_CHECK_ARG(g->x && g->edge_attr && g->src && g->dst && g->in_ptr && g->in_idx && g->atom_ptr,
                "null graph pointer");
  return CGR_OK;
}

size_t max_splitk_floats(const cgr_params_t* p, const cgr_graph_t* g) {
  const int64_t H = p->hidden, E = g->n_bonds, N = g->n_atoms;
  size_t m = 0;
  auto upd = [&](int64_t M_, int64_t N_, int64_t K_) {
    int s = simt_splitk_choose(M_, N_, K_);
    if (s > 1) { size_t f = (size_t)s * M_ * N_; if (f > m) m = f; }
  };
  upd(H, H, E);
  upd(H, p->fa, N);
  upd(H, H, N);
  upd(H, p->fb > 0 ? p->fb : 1, E);
  return m;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// The layer-wise path.  Its GEMMs run either on the SIMT fp32 kernel or -- training on the tcgen05 engine -- on
// tensor cores through tc_train_gemm (operands split to FP16 (hi, lo) on the fly, weights prepared once per step);
// gathers, element-wise kernels and reductions are shared.
// ------------------------------------------------------------------------------------------------
namespace {

struct Opnd {                    // one GEMM operand as stored: [rows, cols] fp32 with row stride ld
  const float* f32 = nullptr;
  int64_t ld = 0;
  bool kmajor = true;            // true: rows = M (or N), cols = K;  false: rows = K, cols = M (or N)
  int wmat = -1;                 // >= 0: matrix of the prepared-weight buffer (tcgen05 path), rows from wrow0
  int64_t wrow0 = 0;
  bool is_x = false;             // data.x (its split is cached per batch)
  bool scaled = false;           // gradient-like operand: needs the amax-based power-of-two scale
};

struct GemmCtx {
  bool tc = false;
  const cgr_params_t* p = nullptr;
  const void* wbuf = nullptr;
  __half *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;     // split scratch slots
  const __half *x_hi = nullptr, *x_lo = nullptr;
  int64_t x_ld = 0;
  unsigned int* amax = nullptr;  // [4]
  float* unscale = nullptr;      // [4]
  int* overflow = nullptr;
  float* partial = nullptr;      // split-K partial sums
  const float* last_a = nullptr; // fp32 tensor currently held by slot A (re-used by consecutive GEMMs)
  bool last_a_scaled = false;
};

int64_t ru64(int64_t v) { return (v + 63) / 64 * 64; }

int make_tc_operand(GemmCtx& c, const Opnd& o, int64_t mn, int64_t K, bool slot_b, TcOperand* out, cudaStream_t st) {
  if (o.wmat >= 0) { *out = tc_weight_operand(c.p, c.wbuf, o.wmat, o.wrow0, !o.kmajor); return CGR_OK; }
  if (o.is_x && c.x_hi) { *out = TcOperand{c.x_hi, c.x_lo, c.x_ld, nullptr, !o.kmajor}; return CGR_OK; }
  const int64_t rows = o.kmajor ? mn : K, cols = o.kmajor ? K : mn;
  __half* hi = slot_b ? c.b_hi : c.a_hi;
  __half* lo = slot_b ? c.b_lo : c.a_lo;
  const int slot = slot_b ? 1 : 0;
  if (slot_b || c.last_a != o.f32 || c.last_a_scaled != o.scaled) {
    int rc = tc_split(o.f32, o.ld, rows, (int)cols, o.scaled, hi, lo, ru64(cols), c.amax + slot, c.unscale + slot,
                      c.overflow, st);
    if (rc) return rc;
    if (!slot_b) { c.last_a = o.f32; c.last_a_scaled = o.scaled; }
  }
  *out = TcOperand{hi, lo, ru64(cols), o.scaled ? c.unscale + slot : nullptr, !o.kmajor};
  return CGR_OK;
}

// C[M,N] = sum_k A(m,k) B(n,k) with the fused epilogue; reduction GEMMs (K >> M,N) use deterministic split-K
int gemm(GemmCtx& c, const Opnd& A, const Opnd& B, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
         const GemmEpilogue& epi, bool allow_splitk, cudaStream_t st) {
  if (!c.tc)
    return simt_gemm(A.f32, A.ld, A.kmajor, B.f32, B.ld, B.kmajor, C, ldc, M, N, K, epi,
                     allow_splitk ? simt_splitk_choose(M, N, K) : 1, c.partial, st);
  TcOperand oa, ob;
  int rc = make_tc_operand(c, A, M, K, false, &oa, st);
  if (rc) return rc;
  rc = make_tc_operand(c, B, N, K, true, &ob, st);
  if (rc) return rc;
  return tc_train_gemm(oa, ob, M, N, K, C, ldc, epi, allow_splitk ? tc_splitk_choose(M, N, K) : 1, c.partial, st);
}

Opnd act_opnd(const float* p, int64_t ld, bool kmajor, bool scaled = false) {
  Opnd o; o.f32 = p; o.ld = ld; o.kmajor = kmajor; o.scaled = scaled; return o;
}
Opnd x_opnd(const float* x, int64_t fa, bool kmajor) {
  Opnd o; o.f32 = x; o.ld = fa; o.kmajor = kmajor; o.is_x = true; return o;
}
Opnd w_opnd(const float* w, int64_t ld, bool kmajor, int wmat, int64_t wrow0 = 0) {
  Opnd o; o.f32 = w; o.ld = ld; o.kmajor = kmajor; o.wmat = wmat; o.wrow0 = wrow0; return o;
}
// without the tcgen05 path the prepared-weight ids are ignored (SIMT reads the fp32 parameter directly)

int edge_init_impl(GemmCtx& c, const float* x, const float* edge_attr, const int32_t* src, const float* w_init,
                   const float* b_init, int64_t N, int64_t E, int fa, int fb, int H, int act, float* h0, float* z0,
                   float* P, cudaStream_t st) {
  // P = x . W_x^T with W_x = w_init[:, :fa]; per-atom projection, algebraically equal to the reference's
  // [x[src] || ea] GEMM (GNN.py:86) without materialising the [E, Fa+Fb] concat.
  GemmEpilogue none;
  none.tag = "gemm_atom_proj";
  int rc = gemm(c, x_opnd(x, fa, true), w_opnd(w_init, fa + fb, true, 0, 0), P, H, N, H, fa, none, false, st);
  if (rc) return rc;
  return simt_edge_init(P, edge_attr, src, w_init, b_init, E, fa, fb, H, act, h0, z0, st);
}

int bond_update_impl(GemmCtx& c, const float* h_in, const float* h0, const int32_t* in_ptr, const int32_t* in_idx,
                     const int32_t* src, const float* w, const float* b, const float* skip, int act, float dropout_p,
                     uint64_t seed, uint32_t layer, int training, float* h_out, float* m_out, float* z_out, int64_t E,
                     int H, int wmat, cudaStream_t st) {
  GatherPost np;
  int rc = simt_gather_bonds(h_in, src, in_ptr, in_idx, 0, m_out, E, H, np, st);   // GNN.py:134-141
  if (rc) return rc;
  c.last_a = nullptr;                                      // m_out may be a buffer re-used every layer: re-split it
  GemmEpilogue e;
  e.bias = b;
  e.res = h0; e.ldr = H; e.res_scale = skip;               // GNN.py:94-97
  e.preact = z_out;
  e.act = act;                                             // GNN.py:100-102
  e.dropout_p = (training && dropout_p > 0.f) ? dropout_p : 0.f;
  e.seed = seed; e.layer = layer;
  e.tag = "gemm_bond_update";
  return gemm(c, act_opnd(m_out, H, true), w_opnd(w, H, true, wmat), h_out, H, E, H, H, e, false, st);
}

int readout_impl(GemmCtx& c, const float* h, const float* x, const int32_t* in_ptr, const int32_t* in_idx,
                 const int32_t* atom_ptr, const float* w_e2n, const float* b_e2n, const float* w_ffn, const float* b_ffn,
                 int act, float* out, float* s_out, float* hv_out, float* zv_out, float* pooled_out, int64_t N, int64_t B,
                 int fa, int H, int depth, cudaStream_t st) {
  int rc = simt_atom_sum(h, in_ptr, in_idx, 0, s_out, N, H, st);                  // GNN.py:105
  if (rc) return rc;
  c.last_a = nullptr;
  // W_o [x || s] = x W_ox^T + s W_os^T  (GNN.py:106-107 without the concat)
  GemmEpilogue e1;
  e1.tag = "gemm_readout_x";
  e1.bias = b_e2n;
  rc = gemm(c, x_opnd(x, fa, true), w_opnd(w_e2n, fa + H, true, 0, H), hv_out, H, N, H, fa, e1, false, st);
  if (rc) return rc;
  GemmEpilogue e2;
  e2.res = hv_out; e2.ldr = H;
  e2.preact = zv_out;
  e2.act = act;
  e2.tag = "gemm_readout_s";
  rc = gemm(c, act_opnd(s_out, H, true), w_opnd(w_e2n + fa, fa + H, true, depth + 1), hv_out, H, N, H, H, e2, false, st);
  if (rc) return rc;
  return simt_pool_ffn(hv_out, atom_ptr, w_ffn, b_ffn, pooled_out, out, B, H, st);  // GNN.py:110
}

// extra workspace of the tcgen05 training path: split slots, scale slots, weights / x split when not supplied
struct TcTrainWs { size_t slot_bytes, w_bytes, x_bytes, total; };
TcTrainWs tc_train_ws(const cgr_params_t* p, const cgr_graph_t* g) {
  TcTrainWs w;
  const int64_t rows = g->n_bonds > g->n_atoms ? g->n_bonds : g->n_atoms;
  w.slot_bytes = cgr_align_up((size_t)rows * ru64(p->hidden) * sizeof(__half), 1024);
  w.w_bytes = p->tc_weights ? 0 : cgr_align_up(tc_weights_bytes(p), 1024);
  w.x_bytes = (g->x_hi && g->x_lo) ? 0 : cgr_align_up((size_t)g->n_atoms * ru64(p->fa) * sizeof(__half), 1024);
  w.total = 4 * w.slot_bytes + w.w_bytes + 2 * w.x_bytes + 1024 + 2048;
  return w;
}

int setup_tc_ctx(GemmCtx& c, const cgr_params_t* p, const cgr_graph_t* g, char* base, cudaStream_t st) {
  const TcTrainWs w = tc_train_ws(p, g);
  char* ptr = (char*)(((uintptr_t)base + 1023) & ~(uintptr_t)1023);
  c.tc = true;
  c.p = p;
  c.a_hi = (__half*)ptr; ptr += w.slot_bytes;
  c.a_lo = (__half*)ptr; ptr += w.slot_bytes;
  c.b_hi = (__half*)ptr; ptr += w.slot_bytes;
  c.b_lo = (__half*)ptr; ptr += w.slot_bytes;
  c.amax = (unsigned int*)ptr;
  c.unscale = (float*)(ptr + 64);
  // fp16-range flag of the operand splits: the caller's tc_status[0] when given (observable), else scratch
  c.overflow = g->tc_status ? g->tc_status : (int*)(ptr + 128);
  ptr += 1024;
  int rc;
  if (p->tc_weights) {
    c.wbuf = p->tc_weights;
  } else {
    rc = tc_prepare_weights(p, ptr, tc_weights_bytes(p), st);
    if (rc) return rc;
    c.wbuf = ptr;
    ptr += w.w_bytes;
  }
  c.x_ld = ru64(p->fa);
  if (g->x_hi && g->x_lo) {
    c.x_hi = (const __half*)g->x_hi; c.x_lo = (const __half*)g->x_lo;
  } else {
    __half* xh = (__half*)ptr; ptr += w.x_bytes;
    __half* xl = (__half*)ptr; ptr += w.x_bytes;
    rc = tc_split_features(g->x, g->n_atoms, p->fa, xh, xl, c.overflow, st);
    if (rc) return rc;
    c.x_hi = xh; c.x_lo = xl;
  }
  return CGR_OK;
}

bool tc_training_ok(const cgr_params_t* p) { return p->depth + 3 <= 16; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// stage-level entry points (SIMT fp32 engine)
// ------------------------------------------------------------------------------------------------

extern "C" int cgr_edge_init_fwd(const float* x, const float* edge_attr, const int32_t* src, const float* w_init,
                                 const float* b_init, int64_t n_atoms, int64_t n_bonds, int32_t fa, int32_t fb,
                                 int32_t hidden, int32_t act, float* h0, float* z0, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  CGR_CHECK_ARG(x && src && w_init && b_init && h0, "cgr_edge_init_fwd: null pointer");
  CGR_CHECK_ARG(fb == 0 || edge_attr, "cgr_edge_init_fwd: edge_attr is null but fb > 0");
  CGR_CHECK_ARG(workspace_bytes >= fbytes((size_t)n_atoms * hidden), "cgr_edge_init_fwd: workspace too small");
  GemmCtx c;
  return edge_init_impl(c, x, edge_attr, src, w_init, b_init, n_atoms, n_bonds, fa, fb, hidden, act, h0, z0,
                        (float*)workspace, (cudaStream_t)stream);
}

extern "C" int cgr_bond_update_fwd(const float* h_in, const float* h0, const int32_t* in_ptr,
                                   const int32_t* in_idx, const int32_t* src, const float* w, const float* b,
                                   const float* skip, int32_t act, float dropout_p, uint64_t seed, uint32_t layer,
                                   int32_t training, float* h_out, float* m_out, float* z_out, int64_t n_bonds,
                                   int64_t n_atoms, int32_t hidden, void* stream) {
  (void)n_atoms;
  CGR_CHECK_ARG(h_in && h0 && in_ptr && in_idx && src && w && b && h_out && m_out, "cgr_bond_update_fwd: null pointer");
  CGR_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "cgr_bond_update_fwd: dropout_p out of range");
  GemmCtx c;
  return bond_update_impl(c, h_in, h0, in_ptr, in_idx, src, w, b, skip, act, dropout_p, seed, layer, training, h_out,
                          m_out, z_out, n_bonds, hidden, -1, (cudaStream_t)stream);
}

extern "C" int cgr_conv_fwd(const float* h, const int32_t* in_ptr, const int32_t* in_idx, const int32_t* src,
                            const float* w, const float* b, float* a_out, float* y_out, float* m_ws,
                            int64_t n_bonds, int64_t n_atoms, int32_t hidden, void* stream) {
  CGR_CHECK_ARG(h && in_ptr && in_idx && src && w && b && a_out && y_out && m_ws, "cgr_conv_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = simt_atom_sum(h, in_ptr, in_idx, 0, a_out, n_atoms, hidden, st);        // GNN.py:134
  if (rc) return rc;
  GatherPost np;
  rc = simt_gather_bonds(h, src, in_ptr, in_idx, 0, m_ws, n_bonds, hidden, np, st); // GNN.py:136-141
  if (rc) return rc;
  GemmEpilogue e;
  e.bias = b;
  return simt_gemm(m_ws, hidden, true, w, hidden, true, y_out, hidden, n_bonds, hidden, hidden, e, 1, nullptr, st);
}

extern "C" int cgr_readout_fwd(const float* h, const float* x, const int32_t* in_ptr, const int32_t* in_idx,
                               const int32_t* atom_ptr, const float* w_e2n, const float* b_e2n, const float* w_ffn,
                               const float* b_ffn, int32_t act, float* out, float* s_out, float* hv_out,
                               float* zv_out, float* pooled_out, int64_t n_atoms, int64_t n_bonds, int64_t n_rxn,
                               int32_t fa, int32_t hidden, void* stream) {
  (void)n_bonds;
  CGR_CHECK_ARG(h && x && in_ptr && in_idx && atom_ptr && w_e2n && b_e2n && w_ffn && b_ffn && out && s_out &&
                    hv_out && pooled_out, "cgr_readout_fwd: null pointer");
  GemmCtx c;
  return readout_impl(c, h, x, in_ptr, in_idx, atom_ptr, w_e2n, b_e2n, w_ffn, b_ffn, act, out, s_out, hv_out, zv_out,
                      pooled_out, n_atoms, n_rxn, fa, hidden, 0, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// stage-level backward entry points (SIMT fp32 engine): the explicit mirror of autograd for each stage
// ------------------------------------------------------------------------------------------------
namespace {
size_t stage_bwd_partial_floats(int64_t H, int64_t E, int64_t N, int64_t fa, int64_t fb) {
  size_t m = 0;
  auto upd = [&](int64_t M_, int64_t N_, int64_t K_) {
    const int s = simt_splitk_choose(M_, N_, K_);
    if (s > 1) { const size_t f = (size_t)s * M_ * N_; if (f > m) m = f; }
  };
  upd(H, H, E); upd(H, fa, N); upd(H, H, N); upd(H, fb > 0 ? fb : 1, E);
  return m;
}
}  // namespace

extern "C" size_t cgr_stage_bwd_workspace(int64_t n_atoms, int64_t n_bonds, int32_t fa, int32_t fb, int32_t hidden) {
  const size_t H = hidden, E = n_bonds, N = n_atoms;
  return 2 * fbytes(N * H) + 2 * fbytes(E * H) + fbytes(stage_bwd_partial_floats(H, E, N, fa, fb)) +
         fbytes(simt_colsum_workspace(E > N ? E : N, (int)H)) + 256;
}

extern "C" int cgr_readout_bwd(const float* grad_out, const float* x, const int32_t* in_ptr, const int32_t* in_idx,
                               const int32_t* atom_ptr, const int32_t* dst, const float* w_e2n, const float* w_ffn,
                               int32_t act, const float* s, const float* hv, const float* zv, const float* pooled,
                               float* gw_e2n, float* gb_e2n, float* gw_ffn, float* gb_ffn, float* dh, int64_t n_atoms,
                               int64_t n_bonds, int64_t n_rxn, int32_t fa, int32_t hidden, void* workspace,
                               size_t workspace_bytes, void* stream) {
  CGR_CHECK_ARG(grad_out && x && in_ptr && in_idx && atom_ptr && dst && w_e2n && w_ffn && s && hv && pooled && gw_e2n &&
                    gb_e2n && gw_ffn && gb_ffn && dh, "cgr_readout_bwd: null pointer");
  CGR_CHECK_ARG(act == CGR_ACT_RELU || zv, "cgr_readout_bwd: zv (pre-activations) required for silu / gelu");
  CGR_CHECK_ARG(workspace_bytes >= cgr_stage_bwd_workspace(n_atoms, n_bonds, fa, 0, hidden), "cgr_readout_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t H = hidden, N = n_atoms, E = n_bonds, B = n_rxn;
  (void)in_ptr; (void)in_idx;
  Carver ws(workspace, workspace_bytes);
  float* dzv = ws.floats(N * H);
  float* ds = ws.floats(N * H);
  float* partial = ws.floats(stage_bwd_partial_floats(H, E, N, fa, 0));
  float* csws = ws.floats(simt_colsum_workspace(E > N ? E : N, (int)H));
  if (!ws.ok) { cgr_set_error("cgr_readout_bwd: workspace carve failed"); return CGR_ERR_WORKSPACE; }
  GemmCtx c;
  c.partial = partial;
  GemmEpilogue none;
  int rc = simt_ffn_grads(grad_out, pooled, gw_ffn, gb_ffn, B, (int)H, st);                       // GNN.py:110
  if (rc) return rc;
  rc = simt_readout_dz(grad_out, atom_ptr, w_ffn, hv, zv, act, dzv, B, N, (int)H, st);            // GNN.py:107
  if (rc) return rc;
  rc = simt_colsum(dzv, N, (int)H, gb_e2n, nullptr, nullptr, nullptr, nullptr, true, csws, st);
  if (rc) return rc;
  rc = gemm(c, act_opnd(dzv, H, false), x_opnd(x, fa, false), gw_e2n, fa + H, H, fa, N, none, true, st);    // dW_o[:, :fa]
  if (rc) return rc;
  rc = gemm(c, act_opnd(dzv, H, false), act_opnd(s, H, false), gw_e2n + fa, fa + H, H, H, N, none, true, st);   // dW_o[:, fa:]
  if (rc) return rc;
  rc = gemm(c, act_opnd(dzv, H, true), act_opnd(w_e2n + fa, fa + H, false), ds, H, N, H, H, none, false, st);  // ds = dzv W_os
  if (rc) return rc;
  GatherPost np;
  return simt_expand_dst(ds, dst, dh, E, (int)H, np, st);                                          // dh_d[e] = ds[dst e]
}

extern "C" int cgr_bond_update_bwd(const float* dh_out, const float* h_out, const float* z, const float* m,
                                   const float* h0, const int32_t* in_ptr, const int32_t* in_idx, const int32_t* dst,
                                   const float* w, const float* skip, int32_t act, float dropout_p, uint64_t seed,
                                   uint32_t layer, int32_t training, float* gw, float* gb, float* gskip, float* dh_in,
                                   float* dh0_acc, int32_t dh0_first, int64_t n_bonds, int64_t n_atoms, int32_t hidden,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  CGR_CHECK_ARG(dh_out && h_out && m && h0 && in_ptr && in_idx && dst && w && gw && gb && dh_in && dh0_acc,
                "cgr_bond_update_bwd: null pointer");
  CGR_CHECK_ARG(act == CGR_ACT_RELU || z, "cgr_bond_update_bwd: z (pre-activations) required for silu / gelu");
  CGR_CHECK_ARG(workspace_bytes >= cgr_stage_bwd_workspace(n_atoms, n_bonds, 1, 0, hidden), "cgr_bond_update_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t H = hidden, E = n_bonds, N = n_atoms;
  Carver ws(workspace, workspace_bytes);
  (void)ws.floats(N * H); (void)ws.floats(N * H);
  float* dz = ws.floats(E * H);
  float* dm = ws.floats(E * H);
  float* partial = ws.floats(stage_bwd_partial_floats(H, E, N, 1, 0));
  float* csws = ws.floats(simt_colsum_workspace(E > N ? E : N, (int)H));
  if (!ws.ok) { cgr_set_error("cgr_bond_update_bwd: workspace carve failed"); return CGR_ERR_WORKSPACE; }
  GemmCtx c;
  c.partial = partial;
  GemmEpilogue none;
  // dz = dh_out (.) act'(z) (.) dropout                                                   (GNN.py:100-102)
  GatherPost gp;
  gp.mode = 1; gp.act = act; gp.h_next = h_out; gp.z = z;
  gp.dropout_p = (training && dropout_p > 0.f) ? dropout_p : 0.f; gp.seed = seed; gp.layer = layer;
  int rc = simt_expand_dst(dh_out, nullptr, dz, E, (int)H, gp, st);
  if (rc) return rc;
  // db = sum dz; dskip = sum dz . h0; dh0 (+)= skip * dz                                    (GNN.py:94-97)
  rc = simt_colsum(dz, E, (int)H, gb, gskip ? h0 : nullptr, gskip, dh0_acc, skip, dh0_first != 0, csws, st);
  if (rc) return rc;
  rc = gemm(c, act_opnd(dz, H, false), act_opnd(m, H, false), gw, H, H, H, E, none, true, st);    // dW = dz^T m
  if (rc) return rc;
  rc = gemm(c, act_opnd(dz, H, true), act_opnd(w, H, false), dm, H, E, H, H, none, false, st);    // dm = dz W
  if (rc) return rc;
  GatherPost np;                                                                                  // GNN.py:134-141 transposed
  return simt_gather_bonds(dm, dst, in_ptr, in_idx, 1, dh_in, E, (int)H, np, st);                  // dh[k] = sum dm[j^1] - dm[k^1]
}

extern "C" int cgr_edge_init_bwd(const float* dh0, const float* h0, const float* z0, const float* x,
                                 const float* edge_attr, const int32_t* in_ptr, const int32_t* in_idx, int32_t act,
                                 float* gw_init, float* gb_init, int64_t n_atoms, int64_t n_bonds, int32_t fa,
                                 int32_t fb, int32_t hidden, void* workspace, size_t workspace_bytes, void* stream) {
  CGR_CHECK_ARG(dh0 && h0 && x && in_ptr && in_idx && gw_init && gb_init, "cgr_edge_init_bwd: null pointer");
  CGR_CHECK_ARG(fb == 0 || edge_attr, "cgr_edge_init_bwd: edge_attr is null but fb > 0");
  CGR_CHECK_ARG(act == CGR_ACT_RELU || z0, "cgr_edge_init_bwd: z0 (pre-activations) required for silu / gelu");
  CGR_CHECK_ARG(workspace_bytes >= cgr_stage_bwd_workspace(n_atoms, n_bonds, fa, fb, hidden), "cgr_edge_init_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t H = hidden, E = n_bonds, N = n_atoms;
  Carver ws(workspace, workspace_bytes);
  float* dP = ws.floats(N * H);
  (void)ws.floats(N * H);
  float* dz = ws.floats(E * H);
  (void)ws.floats(E * H);
  float* partial = ws.floats(stage_bwd_partial_floats(H, E, N, fa, fb));
  float* csws = ws.floats(simt_colsum_workspace(E > N ? E : N, (int)H));
  if (!ws.ok) { cgr_set_error("cgr_edge_init_bwd: workspace carve failed"); return CGR_ERR_WORKSPACE; }
  GemmCtx c;
  c.partial = partial;
  GemmEpilogue none;
  GatherPost gp;                                        // dz0 = dh0 (.) act'(z0)                 (GNN.py:86)
  gp.mode = 1; gp.act = act; gp.h_next = h0; gp.z = z0;
  int rc = simt_expand_dst(dh0, nullptr, dz, E, (int)H, gp, st);
  if (rc) return rc;
  rc = simt_colsum(dz, E, (int)H, gb_init, nullptr, nullptr, nullptr, nullptr, true, csws, st);
  if (rc) return rc;
  if (fb > 0) {                                         // dW_i[:, fa:] = dz0^T ea
    rc = simt_gemm(dz, H, false, edge_attr, fb, false, gw_init + fa, fa + fb, H, fb, E, none,
                   simt_splitk_choose(H, fb, E), partial, st);
    if (rc) return rc;
  }
  // dW_i[:, :fa] = dP^T x with dP[v] = sum_{e: src e = v} dz0[e] = sum_{j in in(v)} dz0[j^1]
  rc = simt_atom_sum(dz, in_ptr, in_idx, 1, dP, N, (int)H, st);
  if (rc) return rc;
  return gemm(c, act_opnd(dP, H, false), x_opnd(x, fa, false), gw_init, fa + fb, H, fa, N, none, true, st);
}

// ------------------------------------------------------------------------------------------------
// whole network
// ------------------------------------------------------------------------------------------------

extern "C" size_t cgr_forward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int32_t training,
                                        int32_t engine) {
  if (!p || !g) return 0;
  // tcgen05 engine: fused tile kernels when a tile plan exists and nothing has to be saved; otherwise the layer-wise
  // path with tensor-core GEMMs (training, or graphs whose reactions exceed a 128-bond tile)
  if (engine == CGR_ENGINE_TC && g->tile_info && g->n_tiles > 0 && (!training || tc_fused_training_ok(p, g))) {
    // training callers may still choose the layer-wise path (no tc_blob): size for the larger of the two
    const size_t f = tc_forward_workspace(p, g, training);
    if (!training) return f;
    const size_t H_ = p->hidden, N_ = g->n_atoms;
    const size_t lw = fbytes(N_ * H_) + tc_train_ws(p, g).total + 256;
    return f > lw ? f : lw;
  }
  const size_t H = p->hidden, E = g->n_bonds, N = g->n_atoms, B = g->n_rxn;
  size_t b = fbytes(N * H);
  if (!training) b += 4 * fbytes(E * H) + 2 * fbytes(N * H) + fbytes(B * H);
  if (engine == CGR_ENGINE_TC) b += tc_train_ws(p, g).total;
  return b + 256;
}

extern "C" size_t cgr_forward_group_workspace(const cgr_params_t* p, const cgr_graph_t* graphs, int32_t n_graphs) {
  if (!p || !graphs) return 0;
  return tc_forward_group_workspace(p, graphs, n_graphs);
}

extern "C" int cgr_gnn_forward_group(const cgr_params_t* p, const cgr_graph_t* graphs, int32_t n_graphs,
                                     float* const* outs, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  CGR_CHECK_ARG(graphs && outs && n_graphs >= 1, "cgr_gnn_forward_group: null argument");
  for (int32_t i = 0; i < n_graphs; ++i)
    if ((rc = check_graph(&graphs[i]))) return rc;
  return tc_gnn_forward_group(p, graphs, n_graphs, outs, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int cgr_gnn_forward(const cgr_params_t* p, const cgr_graph_t* g, float* out, cgr_saved_t* saved,
                               int32_t training, uint64_t seed, int32_t engine, void* workspace,
                               size_t workspace_bytes, void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  rc = check_graph(g);
  if (rc) return rc;
  CGR_CHECK_ARG(out, "cgr_gnn_forward: null output");
  CGR_CHECK_ARG(!training || saved, "cgr_gnn_forward: training forward needs a cgr_saved_t");
  CGR_CHECK_ARG(workspace_bytes >= cgr_forward_workspace(p, g, training, engine),
                "cgr_gnn_forward: workspace too small (%zu < %zu)", workspace_bytes,
                cgr_forward_workspace(p, g, training, engine));
  CGR_CHECK_ARG(engine == CGR_ENGINE_SIMT || engine == CGR_ENGINE_TC, "unknown engine %d", engine);
  // inference on the tcgen05 engine: fused tile kernels.  Training on it: layer-wise path with tensor-core GEMMs.
  if (engine == CGR_ENGINE_TC && (!saved || saved->tc_blob) && g->tile_info && g->n_tiles > 0)
    return tc_gnn_forward(p, g, out, saved, training, seed, workspace, workspace_bytes, (cudaStream_t)stream);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t H = p->hidden, E = g->n_bonds, N = g->n_atoms, B = g->n_rxn;
  const int d = p->depth;
  const size_t EH = (size_t)E * H;
  Carver ws(workspace, workspace_bytes);
  float* P = ws.floats(N * H);
  float *h0, *hbuf[2] = {nullptr, nullptr}, *mbuf = nullptr, *s, *hv, *pooled;
  if (saved) {
    CGR_CHECK_ARG(!saved->tc_blob, "cgr_gnn_forward: tc_blob needs the tcgen05 engine and a tile plan");
    CGR_CHECK_ARG(saved->h_all && saved->m_all && saved->s && saved->hv && saved->pooled,
                  "cgr_gnn_forward: saved buffers missing");
    CGR_CHECK_ARG(p->act == CGR_ACT_RELU || (saved->z_all && saved->zv),
                  "cgr_gnn_forward: z_all/zv required for silu/gelu");
    h0 = saved->h_all; s = saved->s; hv = saved->hv; pooled = saved->pooled;
  } else {
    h0 = ws.floats(EH); hbuf[0] = ws.floats(EH); hbuf[1] = ws.floats(EH); mbuf = ws.floats(EH);
    s = ws.floats(N * H); hv = ws.floats(N * H); pooled = ws.floats(B * H);
  }
  if (!ws.ok) { cgr_set_error("cgr_gnn_forward: workspace carve failed"); return CGR_ERR_WORKSPACE; }
  GemmCtx c;
  if (engine == CGR_ENGINE_TC) {
    CGR_CHECK_ARG(tc_training_ok(p), "tcgen05 training path supports depth <= 13");
    if ((rc = tc_flag_begin(g->tc_status, st))) return rc;
    rc = setup_tc_ctx(c, p, g, ws.base + ws.off, st);
    if (rc) return rc;
  }
  float* z_all = saved ? saved->z_all : nullptr;

  rc = edge_init_impl(c, g->x, g->edge_attr, g->src, p->w_init, p->b_init, N, E, p->fa, p->fb, p->hidden, p->act, h0,
                      z_all, P, st);
  if (rc) return rc;
  const float* h = h0;
  for (int l = 0; l < d; ++l) {
    float* h_out = saved ? saved->h_all + (size_t)(l + 1) * EH : hbuf[l & 1];
    float* m_out = saved ? saved->m_all + (size_t)l * EH : mbuf;
    float* z_out = z_all ? z_all + (size_t)(l + 1) * EH : nullptr;
    const float pdrop = p->host_dropout_p ? p->host_dropout_p[l] : 0.f;
    rc = bond_update_impl(c, h, h0, g->in_ptr, g->in_idx, g->src, p->w_conv[l], p->b_conv[l],
                          p->use_skip ? p->skip[l] : nullptr, p->act, pdrop, seed, (uint32_t)l, training, h_out, m_out,
                          z_out, E, p->hidden, 1 + l, st);
    if (rc) return rc;
    h = h_out;
  }
  rc = readout_impl(c, h, g->x, g->in_ptr, g->in_idx, g->atom_ptr, p->w_e2n, p->b_e2n, p->w_ffn, p->b_ffn, p->act, out,
                    s, hv, saved ? saved->zv : nullptr, pooled, N, B, p->fa, p->hidden, d, st);
  if (rc) return rc;
  if (engine == CGR_ENGINE_TC) return tc_poison_outputs(out, B, g->tc_status, st);
  return CGR_OK;
}

extern "C" size_t cgr_backward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int32_t engine) {
  if (!p || !g) return 0;
  const size_t H = p->hidden, E = g->n_bonds, N = g->n_atoms;
  size_t partial = max_splitk_floats(p, g);
  const size_t tc_partial = (size_t)32 * H * (H > (size_t)p->fa ? H : (size_t)p->fa);
  if (engine == CGR_ENGINE_TC && tc_partial > partial) partial = tc_partial;
  size_t b = 2 * fbytes(N * H) + 3 * fbytes(E * H) + fbytes(partial) +
             fbytes(simt_colsum_workspace(E > N ? E : N, (int)H)) + 256;
  if (engine == CGR_ENGINE_TC) b += tc_train_ws(p, g).total;
  if (engine == CGR_ENGINE_TC) {                // the fused tile-local backward (saved->tc_blob) has its own layout
    const size_t f = tc_backward_workspace(p, g);
    if (f > b) b = f;
  }
  return b;
}

extern "C" int cgr_tc_plan_host(const int64_t* atom_ptr, const int64_t* edge_ptr, int64_t n_rxn, int32_t* tile_info,
                                int64_t* n_tiles) {
  return tc_plan_host(atom_ptr, edge_ptr, n_rxn, tile_info, n_tiles);
}

extern "C" size_t cgr_tc_saved_bytes(const cgr_params_t* p, const cgr_graph_t* g) {
  if (!p || !g) return 0;
  return tc_saved_bytes(p, g);
}

extern "C" int cgr_gnn_backward(const cgr_params_t* p, const cgr_graph_t* g, const cgr_saved_t* saved,
                                const float* grad_out, cgr_grads_t* grads, uint64_t seed, int32_t engine,
                                void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  rc = check_graph(g);
  if (rc) return rc;
  CGR_CHECK_ARG(saved && grad_out && grads, "cgr_gnn_backward: null pointer");
  if (engine == CGR_ENGINE_TC && saved->tc_blob) {      // fused tile-local backward
    CGR_CHECK_ARG(grads->w_init && grads->b_init && grads->w_conv && grads->b_conv && grads->w_e2n && grads->b_e2n &&
                      grads->w_ffn && grads->b_ffn && (!p->use_skip || grads->skip), "cgr_gnn_backward: null grad");
    CGR_CHECK_ARG(workspace_bytes >= tc_backward_workspace(p, g), "cgr_gnn_backward: workspace too small");
    return tc_gnn_backward(p, g, saved, grad_out, grads, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  CGR_CHECK_ARG(saved->h_all && saved->m_all && saved->s && saved->hv && saved->pooled,
                "cgr_gnn_backward: saved buffers missing");
  CGR_CHECK_ARG(p->act == CGR_ACT_RELU || (saved->z_all && saved->zv), "cgr_gnn_backward: z_all/zv required");
  CGR_CHECK_ARG(grads->w_init && grads->b_init && grads->w_conv && grads->b_conv && grads->w_e2n && grads->b_e2n &&
                    grads->w_ffn && grads->b_ffn && (!p->use_skip || grads->skip), "cgr_gnn_backward: null grad");
  CGR_CHECK_ARG(workspace_bytes >= cgr_backward_workspace(p, g, engine), "cgr_gnn_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t H = p->hidden, E = g->n_bonds, N = g->n_atoms, B = g->n_rxn;
  const int d = p->depth, fa = p->fa, fb = p->fb;
  const size_t EH = (size_t)E * H;
  Carver ws(workspace, workspace_bytes);
  float* dzv = ws.floats(N * H);
  float* ds = ws.floats(N * H);
  float* dz = ws.floats(EH);
  float* dm = ws.floats(EH);
  float* dh0 = ws.floats(EH);
  size_t partial_floats = max_splitk_floats(p, g);
  const size_t tc_partial = (size_t)32 * H * (H > fa ? H : fa);
  if (engine == CGR_ENGINE_TC && tc_partial > partial_floats) partial_floats = tc_partial;
  float* partial = ws.floats(partial_floats);
  float* csws = ws.floats(simt_colsum_workspace(E > N ? E : N, (int)H));
  if (!ws.ok) { cgr_set_error("cgr_gnn_backward: workspace carve failed"); return CGR_ERR_WORKSPACE; }
  GemmCtx c;
  if (engine == CGR_ENGINE_TC) {
    CGR_CHECK_ARG(tc_training_ok(p), "tcgen05 training path supports depth <= 13");
    rc = setup_tc_ctx(c, p, g, ws.base + ws.off, st);
    if (rc) return rc;
  }
  c.partial = partial;
  GemmEpilogue none;
  const float* h0 = saved->h_all;

  // ---- readout (GNN.py:105-110) ----
  rc = simt_ffn_grads(grad_out, saved->pooled, grads->w_ffn, grads->b_ffn, B, (int)H, st);
  if (rc) return rc;
  rc = simt_readout_dz(grad_out, g->atom_ptr, p->w_ffn, saved->hv, saved->zv, p->act, dzv, B, N, (int)H, st);
  if (rc) return rc;
  rc = simt_colsum(dzv, N, (int)H, grads->b_e2n, nullptr, nullptr, nullptr, nullptr, true, csws, st);
  if (rc) return rc;
  // dW_o[:, :fa] = dzv^T x ; dW_o[:, fa:] = dzv^T s
  none.tag = "wgrad_readout_x";
  rc = gemm(c, act_opnd(dzv, H, false, true), x_opnd(g->x, fa, false), grads->w_e2n, fa + H, H, fa, N, none, true, st);
  if (rc) return rc;
  none.tag = "wgrad_readout_s";
  rc = gemm(c, act_opnd(dzv, H, false, true), act_opnd(saved->s, H, false), grads->w_e2n + fa, fa + H, H, H, N, none, true,
            st);
  if (rc) return rc;
  // ds = dzv . W_os
  none.tag = "dgrad_readout";
  rc = gemm(c, act_opnd(dzv, H, true, true), w_opnd(p->w_e2n + fa, fa + H, false, d + 1), ds, H, N, H, H, none, false, st);
  if (rc) return rc;

  // ---- message passing layers, last to first ----
  auto post_for_layer = [&](int l) {   // derivative of layer l's activation / dropout
    GatherPost gp;
    gp.mode = 1;
    gp.act = p->act;
    gp.h_next = saved->h_all + (size_t)(l + 1) * EH;
    gp.z = saved->z_all ? saved->z_all + (size_t)(l + 1) * EH : nullptr;
    gp.dropout_p = p->host_dropout_p ? p->host_dropout_p[l] : 0.f;
    gp.seed = seed;
    gp.layer = (uint32_t)l;
    return gp;
  };
  rc = simt_expand_dst(ds, g->dst, dz, E, (int)H, post_for_layer(d - 1), st);   // dh_d[e] = ds[dst e]
  if (rc) return rc;
  for (int l = d - 1; l >= 0; --l) {
    const float* skip = p->use_skip ? p->skip[l] : nullptr;
    // db_l, dskip_l, dh0 += skip_l * dz_l
    rc = simt_colsum(dz, E, (int)H, grads->b_conv[l], p->use_skip ? h0 : nullptr,
                     p->use_skip ? grads->skip[l] : nullptr, dh0, skip, l == d - 1, csws, st);
    if (rc) return rc;
    c.last_a = nullptr;                 // dz was rewritten in place: its split is stale
    // dW_l = dz^T m_l
    none.tag = "wgrad_bond";
    rc = gemm(c, act_opnd(dz, H, false, true), act_opnd(saved->m_all + (size_t)l * EH, H, false), grads->w_conv[l], H, H, H,
              E, none, true, st);
    if (rc) return rc;
    // dm = dz . W_l
    none.tag = "dgrad_bond";
    rc = gemm(c, act_opnd(dz, H, true, true), w_opnd(p->w_conv[l], H, false, 1 + l), dm, H, E, H, H, none, false, st);
    if (rc) return rc;
    // dh_l[k] = sum_{j in in(dst k)} dm[j^1] - dm[k^1]
    GatherPost gp;
    if (l > 0) {
      gp = post_for_layer(l - 1);
    } else {                      // h_0 feeds layer 0 and every skip: dz0 = (dh_0 + dh0) * act'(z_init)
      gp.mode = 2;
      gp.add = dh0;
      gp.act = p->act;
      gp.h_next = h0;
      gp.z = saved->z_all;
      gp.dropout_p = 0.f;
    }
    rc = simt_gather_bonds(dm, g->dst, g->in_ptr, g->in_idx, 1, dz, E, (int)H, gp, st);
    if (rc) return rc;
  }
  c.last_a = nullptr;

  // ---- edge initialisation (GNN.py:85-86) ----
  rc = simt_colsum(dz, E, (int)H, grads->b_init, nullptr, nullptr, nullptr, nullptr, true, csws, st);
  if (rc) return rc;
  if (fb > 0) {       // [H x fb] with fb = 14: too narrow for a tensor-core tile, stays on the fp32 kernel
    GemmEpilogue e;
    e.tag = "wgrad_edge_attr";
    rc = simt_gemm(dz, H, false, g->edge_attr, fb, false, grads->w_init + fa, fa + fb, H, fb, E, e,
                   simt_splitk_choose(H, fb, E), partial, st);
    if (rc) return rc;
  }
  float* dP = ds;   // dP[v] = sum_{e: src e = v} dz0[e] = sum_{j in in(v)} dz0[j^1]
  rc = simt_atom_sum(dz, g->in_ptr, g->in_idx, 1, dP, N, (int)H, st);
  if (rc) return rc;
  none.tag = "wgrad_init_x";
  return gemm(c, act_opnd(dP, H, false, true), x_opnd(g->x, fa, false), grads->w_init, fa + fb, H, fa, N, none, true, st);
}

extern "C" int cgr_tc_plan_build(const int32_t* in_ptr, const int32_t* atom_ptr, int64_t n_rxn, int32_t* tile_info,
                                 int32_t* status, void* stream) {
  return tc_plan_build(in_ptr, atom_ptr, nullptr, nullptr, n_rxn, tile_info, status, (cudaStream_t)stream);
}
extern "C" int cgr_tc_plan_check(const int32_t* tile_info, int64_t n_tiles, const int32_t* src, const int32_t* dst,
                                 int32_t* status, void* stream) {
  CGR_CHECK_ARG(tile_info && src && dst && status, "cgr_tc_plan_check: null pointer");
  return tc_plan_check(tile_info, n_tiles, src, dst, status, (cudaStream_t)stream);
}
extern "C" size_t cgr_tc_gemm_test_workspace(int64_t m, int64_t n, int64_t k) { return tc_gemm2_test_workspace(m, n, k); }
extern "C" int cgr_tc_gemm_test(const float* a, const float* b, int64_t m, int64_t n, int64_t k, int32_t a_mn, int32_t b_mn,
                                float* c, void* workspace, size_t workspace_bytes, void* stream) {
  return tc_gemm2_test(a, b, m, n, k, a_mn, b_mn, c, workspace, workspace_bytes, (cudaStream_t)stream);
}
extern "C" int64_t cgr_tc_features_ld(int32_t fa) { return ((int64_t)fa + 63) / 64 * 64; }
extern "C" int cgr_tc_split_features(const float* x, int64_t n_atoms, int32_t fa, void* x_hi, void* x_lo,
                                     int32_t* status, void* stream) {
  return tc_split_features(x, n_atoms, fa, x_hi, x_lo, status, (cudaStream_t)stream);
}
extern "C" int cgr_tc_debug_buffer(void* p) { tc_set_debug_buffer((long long*)p); return CGR_OK; }
extern "C" size_t cgr_tc_weights_bytes(const cgr_params_t* p) { return p ? tc_weights_bytes(p) : 0; }
extern "C" int cgr_tc_prepare_weights(const cgr_params_t* p, void* buffer, size_t buffer_bytes, void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  return tc_prepare_weights(p, buffer, buffer_bytes, (cudaStream_t)stream);
}
extern "C" size_t cgr_tc_linear_workspace(int64_t m, int64_t n, int64_t k) { return tc_linear_workspace(m, n, k); }
extern "C" int cgr_tc_linear(const float* x, int64_t m, int64_t k, const float* w, int64_t n, const float* bias,
                             float* out, void* workspace, size_t workspace_bytes, void* stream) {
  return tc_linear(x, m, k, k, w, n, k, bias, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int cgr_mse_sum_fwd_bwd(const float* pred, const float* y, int64_t n_rxn, float* loss, float* grad_pred,
                                   void* stream) {
  CGR_CHECK_ARG(pred && y && n_rxn >= 0, "cgr_mse_sum_fwd_bwd: bad argument");
  return simt_mse_sum(pred, y, n_rxn, loss, grad_pred, (cudaStream_t)stream);
}

extern "C" int cgr_dropout_mask(uint64_t seed, uint32_t layer, float dropout_p, int64_t n_bonds, int32_t hidden,
                                uint8_t* mask, void* stream) {
  CGR_CHECK_ARG(mask && n_bonds >= 0 && hidden > 0, "cgr_dropout_mask: bad argument");
  return simt_dropout_mask(seed, layer, dropout_p, n_bonds * hidden, mask, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// End-to-end inference on host buffers
// ------------------------------------------------------------------------------------------------
int csr_by_reaction_shifted(const int64_t* edge_index, const int32_t* edge_ptr, const int32_t* atom_ptr,
                            const int32_t* rxn_shift, int64_t n_rxn, int64_t n_bonds, int64_t n_atoms, int32_t* src,
                            int32_t* dst, int32_t* in_ptr, int32_t* in_idx, int32_t* status, cudaStream_t st);
int store_gather_split(const float* x_all, const float* ea_all, const int32_t* ei_all, const int64_t* node_ptr,
                       const int64_t* edge_ptr, int64_t e_all, const int64_t* sel, const int64_t* out_node_ptr,
                       const int64_t* out_edge_ptr, int64_t n_sel, int32_t fa, int32_t fb, int64_t e_out, void* x_hi,
                       void* x_lo, int64_t ldo, float* edge_attr, int64_t* edge_index, int* range_flag, int flag_bit,
                       cudaStream_t st);
namespace {
struct HostInferLayout {
  int64_t t_max, kp_x;
  size_t o_x, o_ea, o_ei, o_meta, o_src, o_dst, o_inidx, o_inptr, o_status, o_xhi, o_xlo, o_out, o_fwd, dev_total;
  size_t meta_ints, fwd_bytes, host_total;
  size_t stage_bytes;          // [edge_attr | edge_index | meta]: contiguous on the device, mirrored in pinned host memory
};
HostInferLayout host_infer_layout(const cgr_params_t* p, int64_t N, int64_t E, int64_t B) {
  HostInferLayout L;
  L.t_max = 2 * E / 128 + 1 < B ? 2 * E / 128 + 1 : B;       // greedy packing: two consecutive tiles hold > 128 bonds
  if (L.t_max < 1) L.t_max = 1;
  L.kp_x = ((int64_t)p->fa + 63) / 64 * 64;
  L.meta_ints = (size_t)L.t_max * 8 + 2 * (size_t)(B + 1) + (size_t)B;   // tiles, atom_ptr, edge_ptr, atom-id shift
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += cgr_align_up(bytes, 1024); return o; };
  L.o_x = take((size_t)N * p->fa * 4);
  L.o_ea = take((size_t)E * (p->fb > 0 ? p->fb : 1) * 4);
  L.o_ei = take((size_t)2 * E * 8);
  L.o_meta = take(L.meta_ints * 4);
  L.o_src = take((size_t)E * 4);
  L.o_dst = take((size_t)E * 4);
  L.o_inidx = take((size_t)E * 4);
  L.o_inptr = take((size_t)(N + 1) * 4);
  L.o_status = take((size_t)(2 + L.t_max) * 4);
  L.o_xhi = take((size_t)N * L.kp_x * 2);
  L.o_xlo = take((size_t)N * L.kp_x * 2);
  L.o_out = take((size_t)B * 4);
  cgr_graph_t g;
  memset(&g, 0, sizeof(g));
  g.n_atoms = N; g.n_bonds = E; g.n_rxn = B; g.n_tiles = L.t_max;
  g.tile_info = (const int32_t*)16; g.x_hi = (const void*)16; g.x_lo = (const void*)16;   // non-null markers for sizing
  cgr_params_t pp = *p;
  if (!pp.tc_weights) pp.tc_weights = (const void*)16;
  L.fwd_bytes = tc_forward_workspace(&pp, &g, 0);
  L.o_fwd = take(L.fwd_bytes);
  L.dev_total = off + 1024;
  // the small inputs travel as ONE copy: the host workspace mirrors the device span [o_ea, o_meta + meta)
  L.stage_bytes = cgr_align_up(L.o_meta + L.meta_ints * 4 - L.o_ea, 256);
  L.host_total = L.stage_bytes + 256;
  return L;
}
}  // namespace

extern "C" int cgr_infer_host_workspace(const cgr_params_t* p, int64_t n_atoms, int64_t n_bonds, int64_t n_rxn,
                                        size_t* dev_bytes, size_t* host_bytes) {
  CGR_CHECK_ARG(p && dev_bytes && host_bytes && n_atoms > 0 && n_bonds > 0 && n_rxn > 0, "cgr_infer_host_workspace: bad argument");
  const HostInferLayout L = host_infer_layout(p, n_atoms, n_bonds, n_rxn);
  *dev_bytes = L.dev_total;
  *host_bytes = L.host_total;
  return CGR_OK;
}

extern "C" int cgr_gnn_infer_host_multi_async(const cgr_params_t* p, const cgr_host_batch_t* batches, int32_t n_batches,
                                              float* host_out, void* dev_ws, size_t dev_bytes, void* host_ws,
                                              size_t host_bytes, void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  CGR_CHECK_ARG(batches && n_batches > 0 && host_out && dev_ws && host_ws, "cgr_gnn_infer_host: null pointer");
  CGR_CHECK_ARG(p->tc_weights, "cgr_gnn_infer_host: prepare the weights first (cgr_tc_prepare_weights)");
  int64_t N = 0, E = 0, B = 0;
  for (int j = 0; j < n_batches; ++j) {
    const cgr_host_batch_t& hb = batches[j];
    CGR_CHECK_ARG(hb.x && hb.edge_index, "cgr_gnn_infer_host: null pointer in batch %d", j);
    CGR_CHECK_ARG(p->fb == 0 || hb.edge_attr, "cgr_gnn_infer_host: edge_attr missing");
    CGR_CHECK_ARG(hb.n_atoms > 0 && hb.n_bonds > 0 && hb.n_rxn > 0 && (hb.n_bonds & 1) == 0,
                  "cgr_gnn_infer_host: bad sizes");
    CGR_CHECK_ARG(hb.ptr || hb.batch || hb.n_rxn == 1,
                  "cgr_gnn_infer_host: neither ptr nor batch given for a multi-graph batch");
    N += hb.n_atoms; E += hb.n_bonds; B += hb.n_rxn;
  }
  CGR_CHECK_ARG(N < (1ll << 31) && E < (1ll << 31), "cgr_gnn_infer_host: sizes exceed the int32 index range");
  const HostInferLayout L = host_infer_layout(p, N, E, B);
  CGR_CHECK_ARG(dev_bytes >= L.dev_total && host_bytes >= L.host_total, "cgr_gnn_infer_host: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* dws = (char*)(((uintptr_t)dev_ws + 1023) & ~(uintptr_t)1023);
  char* h_stage = (char*)host_ws;                                   // mirrors the device span that starts at o_ea
  float* h_ea = (float*)h_stage;
  int64_t* h_ei = (int64_t*)(h_stage + (L.o_ei - L.o_ea));
  int32_t* h_meta = (int32_t*)(h_stage + (L.o_meta - L.o_ea));
  int32_t* h_tiles = h_meta;
  int32_t* h_aptr = h_meta + (size_t)L.t_max * 8;
  int32_t* h_eptr = h_aptr + (B + 1);
  int32_t* h_shift = h_eptr + (B + 1);
  int32_t* h_flags = (int32_t*)(h_stage + L.stage_bytes);

  // ---- host side: per-reaction offsets of the super-batch (host batches laid end to end) ----
  {
    int64_t a_off = 0, e_off = 0, r_off = 0;
    for (int j = 0; j < n_batches; ++j) {
      const cgr_host_batch_t& hb = batches[j];
      int32_t* ap = h_aptr + r_off;
      if (hb.ptr) {
        for (int64_t g = 0; g < hb.n_rxn; ++g) ap[g] = (int32_t)(a_off + hb.ptr[g]);
      } else if (hb.batch) {
        int64_t v = 0;
        for (int64_t g = 0; g < hb.n_rxn; ++g) {
          ap[g] = (int32_t)(a_off + v);
          while (v < hb.n_atoms && hb.batch[v] <= g) ++v;
        }
      } else {
        ap[0] = (int32_t)a_off;
      }
      ap[hb.n_rxn] = (int32_t)(a_off + hb.n_atoms);
      int64_t e = 0;
      int32_t* ep = h_eptr + r_off;
      ep[0] = (int32_t)e_off;
      for (int64_t g = 0; g < hb.n_rxn; ++g) {          // bonds of a collated batch are grouped by reaction
        const int64_t a1 = ap[g + 1] - a_off;
        while (e < hb.n_bonds && hb.edge_index[e] < a1) ++e;
        ep[g + 1] = (int32_t)(e_off + e);
        h_shift[r_off + g] = (int32_t)a_off;
      }
      if (e != hb.n_bonds) { cgr_set_error("edge_index is not grouped by reaction"); return CGR_ERR_ARG; }
      a_off += hb.n_atoms; e_off += hb.n_bonds; r_off += hb.n_rxn;
    }
  }
  // ---- greedy tile plan (same rule as cgr_tc_plan_build) ----
  int64_t T = -1;
  {
    int used_e = 129, used_a = 129;
    for (int64_t g = 0; g < B; ++g) {
      const int ne = h_eptr[g + 1] - h_eptr[g], na = h_aptr[g + 1] - h_aptr[g];
      if (ne > 128 || na > 128 || ne <= 0 || na <= 0 || (ne & 1)) {
        cgr_set_error("reaction %lld has %d bonds / %d atoms: not tileable for the tcgen05 engine", (long long)g, ne, na);
        return CGR_ERR_UNSUPPORTED;
      }
      if (used_e + ne > 128 || used_a + na > 128) {
        ++T;
        if (T >= L.t_max) { cgr_set_error("tile bound exceeded"); return CGR_ERR_WORKSPACE; }
        int32_t* ti = h_tiles + T * 8;
        ti[0] = h_eptr[g]; ti[1] = 0; ti[2] = h_aptr[g]; ti[3] = 0; ti[4] = (int32_t)g; ti[5] = 0; ti[6] = 0; ti[7] = 0;
        used_e = 0; used_a = 0;
      }
      used_e += ne; used_a += na;
      int32_t* ti = h_tiles + T * 8;
      ti[1] = used_e; ti[3] = used_a; ti[5] += 1;
    }
    ++T;
  }

  // ---- stage inputs, build index arrays, run the forward, fetch the energies ----
  float* d_x = (float*)(dws + L.o_x);
  float* d_ea = (float*)(dws + L.o_ea);
  int64_t* d_ei = (int64_t*)(dws + L.o_ei);
  int32_t* d_meta = (int32_t*)(dws + L.o_meta);
  int32_t* d_status = (int32_t*)(dws + L.o_status);
  float* d_out = (float*)(dws + L.o_out);
  // Atom features (the bulk) are copied straight from each batch's own buffer; bond features, edge_index and the
  // offsets are gathered into the pinned staging area by this host thread and travel as ONE copy (small copies cost
  // the link a few microseconds each, whatever their size).
  {
    int64_t a_off = 0, e_off = 0;
    for (int j = 0; j < n_batches; ++j) {
      const cgr_host_batch_t& hb = batches[j];
      CGR_CUDA(cudaMemcpyAsync(d_x + (size_t)a_off * p->fa, hb.x, (size_t)hb.n_atoms * p->fa * 4,
                               cudaMemcpyHostToDevice, st));
      if (p->fb > 0) memcpy(h_ea + (size_t)e_off * p->fb, hb.edge_attr, (size_t)hb.n_bonds * p->fb * 4);
      memcpy(h_ei + e_off, hb.edge_index, (size_t)hb.n_bonds * 8);                       // row 0 of [2, E_j]
      memcpy(h_ei + E + e_off, hb.edge_index + hb.n_bonds, (size_t)hb.n_bonds * 8);      // row 1
      a_off += hb.n_atoms; e_off += hb.n_bonds;
    }
    CGR_CUDA(cudaMemcpyAsync(dws + L.o_ea, h_stage, L.o_meta + L.meta_ints * 4 - L.o_ea, cudaMemcpyHostToDevice, st));
  }
  CGR_CUDA(cudaMemsetAsync(d_status, 0, (size_t)(2 + L.t_max) * 4, st));
  cgr_graph_t g;
  memset(&g, 0, sizeof(g));
  g.n_atoms = N; g.n_bonds = E; g.n_rxn = B;
  g.x = d_x; g.edge_attr = d_ea;
  g.src = (int32_t*)(dws + L.o_src); g.dst = (int32_t*)(dws + L.o_dst);
  g.in_ptr = (int32_t*)(dws + L.o_inptr); g.in_idx = (int32_t*)(dws + L.o_inidx);
  g.tile_info = d_meta; g.n_tiles = T;
  g.atom_ptr = d_meta + (size_t)L.t_max * 8;
  g.tc_status = d_status + 1;
  g.x_hi = dws + L.o_xhi; g.x_lo = dws + L.o_xlo;
  rc = csr_by_reaction_shifted(d_ei, g.atom_ptr + (B + 1), g.atom_ptr, n_batches > 1 ? g.atom_ptr + 2 * (B + 1) : nullptr,
                               B, E, N, (int32_t*)g.src, (int32_t*)g.dst, (int32_t*)g.in_ptr, (int32_t*)g.in_idx,
                               d_status, st);
  if (rc) return rc;
  rc = tc_split_features(d_x, N, p->fa, (void*)g.x_hi, (void*)g.x_lo, g.tc_status, st);
  if (rc) return rc;
  rc = tc_gnn_forward(p, &g, d_out, nullptr, 0, 0, dws + L.o_fwd, L.fwd_bytes, st);
  if (rc) return rc;
  CGR_CUDA(cudaMemcpyAsync(host_out, d_out, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  CGR_CUDA(cudaMemcpyAsync(h_flags, d_status, 2 * 4, cudaMemcpyDeviceToHost, st));
  return CGR_OK;
}

extern "C" int cgr_gnn_infer_host_async(const cgr_params_t* p, const float* host_x, const float* host_edge_attr,
                                        const int64_t* host_edge_index, const int64_t* host_ptr,
                                        const int64_t* host_batch, int64_t n_atoms, int64_t n_bonds, int64_t n_rxn,
                                        float* host_out, void* dev_ws, size_t dev_bytes, void* host_ws,
                                        size_t host_bytes, void* stream) {
  cgr_host_batch_t hb;
  hb.x = host_x; hb.edge_attr = host_edge_attr; hb.edge_index = host_edge_index; hb.ptr = host_ptr; hb.batch = host_batch;
  hb.n_atoms = n_atoms; hb.n_bonds = n_bonds; hb.n_rxn = n_rxn;
  return cgr_gnn_infer_host_multi_async(p, &hb, 1, host_out, dev_ws, dev_bytes, host_ws, host_bytes, stream);
}

// ------------------------------------------------------------------------------------------------
// Inference over a device-resident reaction store: the per-batch loop in C, pipelined over
// streams.  Per batch: offsets + tile plan on the host, one small upload, gather kernel, one-launch CSR, feature
// split, the two forward launches writing straight into the caller's result vector.
// ------------------------------------------------------------------------------------------------
namespace {
struct StoreBatchSizes { int64_t n_max, e_max, b_max; };
int store_batch_sizes(const cgr_store_t* s, const int64_t* order, int64_t n_total, int64_t bs, StoreBatchSizes* out) {
  int64_t n_max = 0, e_max = 0;
  for (int64_t lo = 0; lo < n_total; lo += bs) {
    const int64_t hi = lo + bs < n_total ? lo + bs : n_total;
    int64_t n = 0, e = 0;
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t r = order[i];
      CGR_CHECK_ARG(r >= 0 && r < s->n_rxn, "cgr_store_infer: reaction id %lld out of range", (long long)r);
      n += s->node_ptr_host[r + 1] - s->node_ptr_host[r];
      e += s->edge_ptr_host[r + 1] - s->edge_ptr_host[r];
    }
    if (n > n_max) n_max = n;
    if (e > e_max) e_max = e;
  }
  out->n_max = n_max; out->e_max = e_max; out->b_max = bs < n_total ? bs : n_total;
  return CGR_OK;
}
// Reactions are independent and a reaction's energy does not depend on which others share its batch (tiles hold whole
// reactions, every reduction is per reaction), so the screening loop assembles SUPER-batches: consecutive batches of the
// caller's size up to ~1024 reactions go through one gather / CSR / split / forward (fewer, larger launches:
// batch 64 screens at the rate of batch 1024).  The result vector is the same, bit for bit.
int64_t store_effective_batch(int64_t batch_size) {
  static const bool off = getenv("CGR_STORE_NO_COALESCE") != nullptr;      // experiments
  if (off || batch_size >= 1024) return batch_size;
  return (1024 / batch_size) * batch_size;
}
// host staging of one slot: [sel | out_node_ptr | out_edge_ptr | perm] (int64) then the int32 meta block of the host entry
size_t store_sel_bytes(int64_t b_max) { return cgr_align_up((size_t)(4 * b_max + 2) * 8, 1024); }

// The screening loop OWNS the order of the reactions inside a batch (the caller only sees the result vector), and the
// tile plan packs CONSECUTIVE whole reactions into 128-row tiles: in the caller's order a T1x-shaped batch fills its
// tiles to ~83 % (half a reaction is lost per tile on average).  Best-fit decreasing on the directed-bond counts --
// bucketed, the counts are even integers <= 128 -- under the second constraint (atoms <= 128) orders the batch so that
// the greedy plan closes every tile nearly full (~96 %): ~13 % fewer tiles for every kernel of the forward.  A
// reaction's energy does not depend on its position (tiles hold whole reactions, every reduction is per reaction), so
// the result vector is the same bit for bit; the energies are scattered back to the caller's order by one small kernel.
// perm[i] = position in the caller's batch of the reaction assembled at position i.  Returns false (identity order)
// when a reaction is not tileable.
bool store_pack_order(const cgr_store_t* s, const int64_t* ord, int64_t B, std::vector<int32_t>& perm,
                      std::vector<int32_t>& nxt) {
  constexpr int TMAX = 128, NB = TMAX / 2;
  int32_t head[NB + 1];
  for (int c = 0; c <= NB; ++c) head[c] = -1;
  perm.resize((size_t)B);
  nxt.resize((size_t)B);
  for (int64_t i = B - 1; i >= 0; --i) {                 // lists end up in ascending batch position
    const int64_t r = ord[i];
    const int64_t ne = s->edge_ptr_host[r + 1] - s->edge_ptr_host[r], na = s->node_ptr_host[r + 1] - s->node_ptr_host[r];
    if (ne <= 0 || ne > TMAX || (ne & 1) || na <= 0 || na > TMAX) return false;
    nxt[(size_t)i] = head[ne >> 1];
    head[ne >> 1] = (int32_t)i;
  }
  auto atoms = [&](int32_t i) { const int64_t r = ord[i]; return (int)(s->node_ptr_host[r + 1] - s->node_ptr_host[r]); };
  int64_t placed = 0;
  int top = NB;
  while (placed < B) {
    while (top > 0 && head[top] < 0) --top;              // largest reaction left opens the tile
    int cap_e = TMAX, cap_a = TMAX;
    int c = top;
    while (c > 0) {
      const int32_t i = head[c];
      if (i < 0 || 2 * c > cap_e || atoms(i) > cap_a) { --c; continue; }
      head[c] = nxt[(size_t)i];
      perm[(size_t)placed++] = i;
      cap_e -= 2 * c;
      cap_a -= atoms(i);
      if (c > cap_e / 2) c = cap_e / 2;
    }
  }
  return true;
}

__global__ void store_scatter_out_kernel(const float* __restrict__ tmp, const int64_t* __restrict__ perm, float* __restrict__ out,
                                         int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[perm[i]] = tmp[i];
}
}  // namespace

// The batch order of the screening loop as a host function of its own (pure host code): the loop's packing is testable
// without a device, and a caller that assembles batches itself (ReactionStore.batch) can ask for the same order.
extern "C" int cgr_store_pack_order(const int64_t* node_ptr_host, const int64_t* edge_ptr_host, int64_t n_rxn_store,
                                    const int64_t* ids, int64_t n, int32_t* perm) {
  CGR_CHECK_ARG(node_ptr_host && edge_ptr_host && ids && perm && n > 0 && n_rxn_store > 0, "cgr_store_pack_order: bad argument");
  for (int64_t i = 0; i < n; ++i) {
    CGR_CHECK_ARG(ids[i] >= 0 && ids[i] < n_rxn_store, "cgr_store_pack_order: reaction id %lld out of range", (long long)ids[i]);
    perm[i] = (int32_t)i;
  }
  cgr_store_t s;
  memset(&s, 0, sizeof(s));
  s.node_ptr_host = node_ptr_host; s.edge_ptr_host = edge_ptr_host; s.n_rxn = n_rxn_store;
  std::vector<int32_t> p, nxt;
  if (!store_pack_order(&s, ids, n, p, nxt)) {
    cgr_set_error("cgr_store_pack_order: a reaction does not fit a 128-row tile (identity order returned)");
    return CGR_ERR_UNSUPPORTED;
  }
  memcpy(perm, p.data(), (size_t)n * sizeof(int32_t));
  return CGR_OK;
}

extern "C" int cgr_store_infer_workspace(const cgr_params_t* p, const cgr_store_t* store, const int64_t* order,
                                         int64_t n_total, int64_t batch_size, size_t* dev_bytes_per_slot,
                                         size_t* host_bytes_per_slot) {
  CGR_CHECK_ARG(p && store && order && n_total > 0 && batch_size > 0 && dev_bytes_per_slot && host_bytes_per_slot,
                "cgr_store_infer_workspace: bad argument");
  batch_size = store_effective_batch(batch_size);
  StoreBatchSizes z;
  int rc = store_batch_sizes(store, order, n_total, batch_size, &z);
  if (rc) return rc;
  const HostInferLayout L = host_infer_layout(p, z.n_max, z.e_max, z.b_max);
  *dev_bytes_per_slot = cgr_align_up(L.dev_total + store_sel_bytes(z.b_max) + 1024, 1024);
  *host_bytes_per_slot = cgr_align_up(store_sel_bytes(z.b_max) + L.meta_ints * 4 + 1024, 1024);
  return CGR_OK;
}

extern "C" int cgr_store_infer(const cgr_params_t* p, const cgr_store_t* store, const int64_t* order, int64_t n_total,
                               int64_t batch_size, float* out, void* dev_ws, size_t dev_bytes_per_slot, void* host_ws,
                               size_t host_bytes_per_slot, int32_t n_slots, void* const* streams) {
  int rc = check_params(p);
  if (rc) return rc;
  CGR_CHECK_ARG(store && order && out && dev_ws && host_ws && streams && n_slots > 0 && n_total > 0 && batch_size > 0,
                "cgr_store_infer: bad argument");
  CGR_CHECK_ARG(p->tc_weights, "cgr_store_infer: prepare the weights first (cgr_tc_prepare_weights)");
  CGR_CHECK_ARG(store->fa == p->fa && store->fb == p->fb, "cgr_store_infer: feature widths of store and model differ");
  batch_size = store_effective_batch(batch_size);
  StoreBatchSizes z;
  if ((rc = store_batch_sizes(store, order, n_total, batch_size, &z))) return rc;
  const HostInferLayout Lmax = host_infer_layout(p, z.n_max, z.e_max, z.b_max);
  const size_t sel_bytes = store_sel_bytes(z.b_max);
  CGR_CHECK_ARG(dev_bytes_per_slot >= Lmax.dev_total + sel_bytes + 1024 &&
                    host_bytes_per_slot >= sel_bytes + Lmax.meta_ints * 4 + 1024, "cgr_store_infer: workspace too small");
  std::vector<cudaEvent_t> ev((size_t)n_slots, nullptr);
  std::vector<char> used((size_t)n_slots, 0);
  auto cleanup = [&]() { for (auto e : ev) if (e) cudaEventDestroy(e); };
  for (int s = 0; s < n_slots; ++s)
    if (cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming) != cudaSuccess) { cleanup(); cgr_set_error("cudaEventCreate failed"); return CGR_ERR_ARG; }
  cgr_params_t pp = *p;
  pp.tc_throughput = 1;                      // several batches in flight
  static const bool pack = getenv("CGR_STORE_NO_PACK") == nullptr;      // experiments: keep the caller's order
  std::vector<int32_t> perm, perm_next;
  const HostInferLayout& L = Lmax;           // one layout for every batch: fixed offsets, sizes vary
  // status words of a slot: [0] validity flags and [1] fp16-range flag are sticky (atomicOr), the readout counters behind
  // them reset themselves -- cleared once, read back once
  rc = CGR_OK;
  for (int s = 0; s < n_slots && rc == CGR_OK; ++s) {
    char* dbase = (char*)(((uintptr_t)dev_ws + (size_t)s * dev_bytes_per_slot + 1023) & ~(uintptr_t)1023);
    const cudaError_t e = cudaMemsetAsync(dbase + sel_bytes + L.o_status, 0, (size_t)(2 + L.t_max) * 4, (cudaStream_t)streams[s]);
    if (e != cudaSuccess) { cgr_set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  int64_t it = 0;
  for (int64_t lo = 0; lo < n_total && rc == CGR_OK; lo += batch_size, ++it) {
    const int slot = (int)(it % n_slots);
    cudaStream_t st = (cudaStream_t)streams[slot];
    const int64_t B = lo + batch_size < n_total ? batch_size : n_total - lo;
    char* hws = (char*)host_ws + (size_t)slot * host_bytes_per_slot;
    char* dbase = (char*)(((uintptr_t)dev_ws + (size_t)slot * dev_bytes_per_slot + 1023) & ~(uintptr_t)1023);
    // the previous batch of this slot must have consumed the pinned staging area
    if (used[slot] && cudaEventSynchronize(ev[slot]) != cudaSuccess) { rc = CGR_ERR_ARG; cgr_set_error("event sync failed"); break; }
    int64_t* h_sel = (int64_t*)hws;
    int64_t* h_optr = h_sel + B;
    int64_t* h_oeptr = h_optr + (B + 1);
    int64_t* h_perm = h_oeptr + (B + 1);
    h_optr[0] = 0; h_oeptr[0] = 0;
    const bool packed = pack && store_pack_order(store, order + lo, B, perm, perm_next);
    for (int64_t i = 0; i < B; ++i) {
      const int64_t pi = packed ? perm[(size_t)i] : i;
      const int64_t r = order[lo + pi];
      h_perm[i] = pi;
      h_sel[i] = r;
      h_optr[i + 1] = h_optr[i] + (store->node_ptr_host[r + 1] - store->node_ptr_host[r]);
      h_oeptr[i + 1] = h_oeptr[i] + (store->edge_ptr_host[r + 1] - store->edge_ptr_host[r]);
    }
    const int64_t N = h_optr[B], E = h_oeptr[B];
    int32_t* h_meta = (int32_t*)(hws + sel_bytes);
    int32_t* h_tiles = h_meta;
    int32_t* h_aptr = h_meta + (size_t)L.t_max * 8;
    int32_t* h_eptr = h_aptr + (B + 1);
    for (int64_t i = 0; i <= B; ++i) { h_aptr[i] = (int32_t)h_optr[i]; h_eptr[i] = (int32_t)h_oeptr[i]; }
    int64_t T = 0;
    rc = tc_plan_host(h_optr, h_oeptr, B, h_tiles, &T);
    if (rc) break;
    if (T > L.t_max) { cgr_set_error("tile bound exceeded"); rc = CGR_ERR_WORKSPACE; break; }
    // device slot: [sel block | host-entry layout]
    int64_t* d_sel = (int64_t*)dbase;
    char* dws = dbase + sel_bytes;
    float* d_x = (float*)(dws + L.o_x);
    float* d_ea = (float*)(dws + L.o_ea);
    int64_t* d_ei = (int64_t*)(dws + L.o_ei);
    int32_t* d_meta = (int32_t*)(dws + L.o_meta);
    int32_t* d_status = (int32_t*)(dws + L.o_status);
    auto CK = [&](cudaError_t e, const char* what) {
      if (e != cudaSuccess && rc == CGR_OK) { cgr_set_error("%s failed: %s", what, cudaGetErrorString(e)); rc = (int)e; }
    };
    CK(cudaMemcpyAsync(d_sel, h_sel, (size_t)(4 * B + 2) * 8, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(sel)");
    CK(cudaMemcpyAsync(d_meta, h_meta, ((size_t)L.t_max * 8 + 2 * (size_t)(B + 1)) * 4, cudaMemcpyHostToDevice, st),
       "cudaMemcpyAsync(meta)");
    CK(cudaEventRecord(ev[slot], st), "cudaEventRecord");
    used[slot] = 1;
    if (rc) break;
    // gather + feature split in one pass: the atom features go straight from the store rows to the FP16 (hi, lo)
    // operand pair of the atom projection (no fp32 copy of x in the slot); 2 = FLAG_X of tc_status[0] (tc.cu)
    rc = store_gather_split(store->x_all, store->ea_all, store->ei_all, store->node_ptr, store->edge_ptr, store->e_all,
                            d_sel, d_sel + B, d_sel + 2 * B + 1, B, p->fa, p->fb, E, dws + L.o_xhi, dws + L.o_xlo, L.kp_x,
                            d_ea, d_ei, d_status + 1, 2, st);
    if (rc) break;
    cgr_graph_t g;
    memset(&g, 0, sizeof(g));
    g.n_atoms = N; g.n_bonds = E; g.n_rxn = B;
    g.x = d_x; g.edge_attr = d_ea;
    g.src = (int32_t*)(dws + L.o_src); g.dst = (int32_t*)(dws + L.o_dst);
    g.in_ptr = (int32_t*)(dws + L.o_inptr); g.in_idx = (int32_t*)(dws + L.o_inidx);
    g.tile_info = d_meta; g.n_tiles = T;
    g.atom_ptr = d_meta + (size_t)L.t_max * 8;
    g.tc_status = d_status + 1;
    g.x_hi = dws + L.o_xhi; g.x_lo = dws + L.o_xlo;
    rc = csr_by_reaction_shifted(d_ei, g.atom_ptr + (B + 1), g.atom_ptr, nullptr, B, E, N, (int32_t*)g.src, (int32_t*)g.dst,
                                 (int32_t*)g.in_ptr, (int32_t*)g.in_idx, d_status, st);
    if (rc) break;
    // packed order: energies land in the slot's own vector and are scattered to the caller's positions
    float* d_tmp = (float*)(dws + L.o_out);
    rc = tc_gnn_forward(&pp, &g, packed ? d_tmp : out + lo, nullptr, 0, 0, dws + L.o_fwd, L.fwd_bytes, st);
    if (rc) break;
    if (packed) {
      cgr_note_launch("store_scatter_out", st, 1);
      store_scatter_out_kernel<<<(unsigned)cgr_ceil_div(B, 256), 256, 0, st>>>(d_tmp, d_sel + 3 * B + 2, out + lo, B);
      CK(cudaGetLastError(), "store_scatter_out_kernel");
    }
  }
  // validity / range flags of every slot
  for (int s = 0; s < n_slots && rc == CGR_OK; ++s) {
    char* dbase = (char*)(((uintptr_t)dev_ws + (size_t)s * dev_bytes_per_slot + 1023) & ~(uintptr_t)1023);
    int32_t* h_flags = (int32_t*)((char*)host_ws + (size_t)s * host_bytes_per_slot + host_bytes_per_slot - 256);
    const cudaError_t e = cudaMemcpyAsync(h_flags, dbase + sel_bytes + L.o_status, 8, cudaMemcpyDeviceToHost, (cudaStream_t)streams[s]);
    if (e != cudaSuccess) { cgr_set_error("cudaMemcpyAsync(flags) failed: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  for (int s = 0; s < n_slots; ++s) {
    const cudaError_t e = cudaStreamSynchronize((cudaStream_t)streams[s]);
    if (e != cudaSuccess && rc == CGR_OK) { cgr_set_error("stream sync failed: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  cleanup();
  if (rc == CGR_OK) {
    int f0 = 0, f1 = 0;
    for (int s = 0; s < n_slots; ++s) {
      const int32_t* h_flags = (const int32_t*)((const char*)host_ws + (size_t)s * host_bytes_per_slot + host_bytes_per_slot - 256);
      f0 |= h_flags[0]; f1 |= h_flags[1];
    }
    if (f0 & 1) { cgr_set_error("directed bonds are not adjacent (e, e^1) reverse pairs"); return CGR_ERR_ARG; }
    if (f0 & 2) { cgr_set_error("a bond leaves its reaction's atom range"); return CGR_ERR_ARG; }
    if (f0 & 4) { cgr_set_error("an atom has no incoming bond (reference GNN.py:106 raises on this input)"); return CGR_ERR_ARG; }
    if (f1) { cgr_set_error("an activation exceeded the fp16 range of the FP16x3 split"); return CGR_ERR_UNSUPPORTED; }
  }
  return rc;
}

extern "C" int cgr_infer_host_check(const cgr_params_t* p, int64_t n_atoms, int64_t n_bonds, int64_t n_rxn,
                                    const void* host_ws) {
  CGR_CHECK_ARG(p && host_ws, "cgr_infer_host_check: null pointer");
  const HostInferLayout L = host_infer_layout(p, n_atoms, n_bonds, n_rxn);
  const int32_t* h_flags = (const int32_t*)((const char*)host_ws + L.stage_bytes);
  if (h_flags[0] & 1) { cgr_set_error("directed bonds are not adjacent (e, e^1) reverse pairs"); return CGR_ERR_ARG; }
  if (h_flags[0] & 2) { cgr_set_error("a bond leaves its reaction's atom range"); return CGR_ERR_ARG; }
  if (h_flags[0] & 4) { cgr_set_error("an atom has no incoming bond (reference GNN.py:106 raises on this input)"); return CGR_ERR_ARG; }
  if (h_flags[1]) { cgr_set_error("an activation exceeded the fp16 range of the FP16x3 split"); return CGR_ERR_UNSUPPORTED; }
  return CGR_OK;
}

extern "C" int cgr_gnn_infer_host(const cgr_params_t* p, const float* host_x, const float* host_edge_attr,
                                  const int64_t* host_edge_index, const int64_t* host_ptr, const int64_t* host_batch,
                                  int64_t n_atoms, int64_t n_bonds, int64_t n_rxn, float* host_out, void* dev_ws,
                                  size_t dev_bytes, void* host_ws, size_t host_bytes, void* stream) {
  int rc = cgr_gnn_infer_host_async(p, host_x, host_edge_attr, host_edge_index, host_ptr, host_batch, n_atoms, n_bonds,
                                    n_rxn, host_out, dev_ws, dev_bytes, host_ws, host_bytes, stream);
  if (rc) return rc;
  CGR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return cgr_infer_host_check(p, n_atoms, n_bonds, n_rxn, host_ws);
}
