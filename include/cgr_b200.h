/*
 * libcgr_b200 — C ABI of the B200-native CGR-MPNN-3D hot path.
 *
 * Drop-in boundary for the reference model `cgr_mpnn_3D/models/GNN.py` (GNN.forward :76-110,
 * DMPNNConv.forward :131-145) and for the batch collation its loaders perform
 * (`cgr_mpnn_3D/training/trainer.py:105-118`, `test.py:85-90`).  The reference has no FFI of
 * its own (it is pure Python on ATen + torch_geometric); these entry points are what a ctypes
 * binding on the reference side calls instead of those Python lines.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the calling thread's current device unless the
 *     parameter name starts with `host_`;
 *   - nothing is allocated or retained: the caller owns outputs and workspaces;
 *   - `stream` is a `cudaStream_t` passed as `void*`;
 *   - return 0 on success, a positive `cudaError_t` value, or a negative CGR_ERR_* code;
 *     `cgr_last_error_string()` describes the last failure on the calling thread;
 *   - functions are re-entrant (autograd runs backward on a different thread).
 *   - all floating tensors are fp32 row-major, index tensors are int32 unless stated.
 */
#ifndef CGR_B200_H_
#define CGR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGR_B200_VERSION 100

/* activation ids: reference train.py:284-292 maps ReLU/SiLU/GELU to F.relu/F.silu/F.gelu */
#define CGR_ACT_RELU_ID 0
#define CGR_ACT_SILU_ID 1
#define CGR_ACT_GELU_ID 2

/* engines: 0 = SIMT fp32 (exact fp32 FMA, layer-wise), 1 = tcgen05 FP16x3 split (tensor cores) */
#define CGR_ENGINE_SIMT 0
#define CGR_ENGINE_TC 1

int cgr_version(void);
const char* cgr_last_error_string(void);

/* ------------------------------------------------------------------------------------------
 * Parameter tree of the reference model (GNN.py:53-74), same tensors as its state_dict:
 *   edge_init.{weight [H,Fa+Fb], bias [H]}, convs.{l}.lin.{weight [H,H], bias [H]},
 *   edge_to_node.{weight [H,Fa+H], bias [H]}, ffn.{weight [1,H], bias [1]}, skip_weights.{l} [].
 * `host_dropout_p`, `w_conv`, `b_conv`, `skip` are HOST arrays of `depth` entries
 * (device pointers inside).  `skip == NULL` <=> use_learnable_skip=False (weight 1).
 * ---------------------------------------------------------------------------------------- */
typedef struct cgr_params {
  int32_t fa, fb, hidden, depth, act, use_skip;
  const float* w_init;
  const float* b_init;
  const float* const* w_conv;
  const float* const* b_conv;
  const float* const* skip;
  const float* w_e2n;
  const float* b_e2n;
  const float* w_ffn;
  const float* b_ffn;
  const float* host_dropout_p;
  const void* tc_weights;    /* optional: buffer filled by cgr_tc_prepare_weights (NULL: prepared per call) */
  int32_t tc_throughput;     /* tcgen05 engine: 1 = the caller pipelines several forwards over streams, prefer the
                                wide-slice / two-tiles-per-cluster configuration; 0 = optimise the latency of a lone forward */
  int32_t tc_fast;           /* tcgen05 engine, inference: 1 = "fast" precision mode -- single-pass fp16 operands (one MMA
                                per k-step instead of the three of the FP16x3 split), ~1e-3 instead of 2e-6 of fp64; NOT
                                the parity mode, reported separately.  0 = fp32-parity mode (default) */
} cgr_params_t;

/* gradient buffers, same shapes as cgr_params (written, not accumulated) */
typedef struct cgr_grads {
  float* w_init;
  float* b_init;
  float* const* w_conv;
  float* const* b_conv;
  float* const* skip;
  float* w_e2n;
  float* b_e2n;
  float* w_ffn;
  float* b_ffn;
} cgr_grads_t;

/* A collated batch of CGR graphs plus the derived index arrays (cgr_csr_build). */
typedef struct cgr_graph {
  int64_t n_atoms, n_bonds, n_rxn;
  const float* x;            /* [N, Fa]  data.x (atom features, MACE block included) */
  const float* edge_attr;    /* [E, Fb]  data.edge_attr */
  const int32_t* src;        /* [E]   b2a   = edge_index[0]           (GNN.py:85)  */
  const int32_t* dst;        /* [E]         = edge_index[1]                         */
  const int32_t* in_ptr;     /* [N+1] a2b CSR offsets of bonds grouped by dst (GNN.py:134) */
  const int32_t* in_idx;     /* [E]   a2b CSR bond ids, ascending inside a group    */
  const int32_t* atom_ptr;   /* [B+1] first atom of each reaction (Batch.ptr)       */
  const int32_t* tile_info;  /* [n_tiles, 8] tile plan of the tcgen05 engine (cgr_tc_plan_build) or NULL */
  int64_t n_tiles;
  int32_t* tc_status;        /* tcgen05 engine: [1 + n_tiles] ints, zero-initialised once by the caller:
                                [0] fp16-range flags of the FP16x3 split -- bit 0: an activation of the LAST forward
                                left the range (cleared when a forward starts), bit 1: a feature of data.x did (set by
                                cgr_tc_split_features, kept for the batch's life); whenever a bit is set the forward
                                writes NaN energies; [1..] self-resetting readout counters.  The layer-wise path (no
                                tile plan) only uses [0] and accepts a 1-int array. */
  const void* x_hi;          /* optional: data.x as FP16 (hi, lo) rows prepared by cgr_tc_split_features */
  const void* x_lo;          /*           (row stride cgr_tc_features_ld(fa) halfs); NULL: converted per call */
} cgr_graph_t;

/* activations kept for the backward pass (caller-allocated; NULL members are not written) */
typedef struct cgr_saved {
  float* h_all;     /* [(depth+1), E, H]  h_0 .. h_depth (post activation / dropout)      */
  float* m_all;     /* [depth, E, H]      gathered messages m_l = a[src] - h[rev]          */
  float* z_all;     /* [(depth+1), E, H]  pre-activations, only for act != relu, else NULL */
  float* s;         /* [N, H]  bond->atom sums of h_depth (GNN.py:105)                     */
  float* hv;        /* [N, H]  atom hidden states (GNN.py:107)                             */
  float* zv;        /* [N, H]  their pre-activations, only for act != relu, else NULL      */
  float* pooled;    /* [B, H]  per-reaction sums (GNN.py:110)                              */
  void* tc_blob;    /* tcgen05 fused training path: ONE buffer of cgr_tc_saved_bytes() bytes that replaces all of
                       the above (they may be NULL): per-layer FP16 (hi, lo) operands, h_0 and hv.  NULL selects
                       the layer-wise path. */
  size_t tc_blob_bytes;
} cgr_saved_t;

/* ------------------------------------------------------------------------------------------
 * (1) Collation — replaces PyG Batch.from_data_list at trainer.py:105-118 / test.py:85-90.
 * `local_edge_index` is the concatenation over graphs of the per-graph [2,e_g] int64 arrays
 * (row 0 = all sources, row 1 = all targets, each of total length E).  Outputs, all int64 like
 * the reference: edge_index [2,E] with cumulative node offsets, batch [N], ptr [B+1],
 * edge_ptr [B+1].  workspace >= cgr_collate_workspace(B).
 * ---------------------------------------------------------------------------------------- */
size_t cgr_collate_workspace(int64_t n_rxn);
int cgr_collate_indices(const int64_t* n_nodes, const int64_t* n_edges,
                        const int64_t* local_edge_index, int64_t n_rxn, int64_t n_bonds,
                        int64_t n_atoms, int64_t* edge_index, int64_t* batch, int64_t* ptr,
                        int64_t* edge_ptr, void* workspace, size_t workspace_bytes, void* stream);

/* CSR a2b / b2a / b2revb from a batched edge_index [2,E] int64 (GNN.py:85,132-138).
 * status[0] bit0: E odd or some bond e^1 is not the reverse of e; bit1: an index is out of
 * range; bit2: an atom without incoming bond (reference raises at GNN.py:106).
 * workspace >= cgr_csr_workspace(N, E). */
size_t cgr_csr_workspace(int64_t n_atoms, int64_t n_bonds);
int cgr_csr_build(const int64_t* edge_index, int64_t n_bonds, int64_t n_atoms, int32_t* src,
                  int32_t* dst, int32_t* in_ptr, int32_t* in_idx, int32_t* status,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Same arrays in ONE launch (one block per reaction) when the per-reaction offsets edge_ptr / atom_ptr
 * (int32 [B+1]) are known and every reaction has <= 256 atoms and bonds; `status` must be zeroed by the
 * caller (bit 3 = a reaction exceeds the in-kernel limit). */
int cgr_csr_build_by_reaction(const int64_t* edge_index, const int32_t* edge_ptr, const int32_t* atom_ptr,
                              int64_t n_rxn, int64_t n_bonds, int64_t n_atoms, int32_t* src, int32_t* dst,
                              int32_t* in_ptr, int32_t* in_idx, int32_t* status, void* stream);

/* atom_ptr [B+1] (int32) from a sorted `batch` vector [N] int64 (Batch.batch). */
int cgr_atom_ptr_from_batch(const int64_t* batch, int64_t n_atoms, int64_t n_rxn,
                            int32_t* atom_ptr, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2)-(4) Stage-level forward entry points (one per north-star subsystem; their backward: (5) below).
 * ---------------------------------------------------------------------------------------- */

/* Edge initialisation, GNN.py:85-86: h0[e] = act(W_i [x[src e] || ea[e]] + b_i).
 * z0 (pre-activation) may be NULL.  workspace >= N*H floats. */
int cgr_edge_init_fwd(const float* x, const float* edge_attr, const int32_t* src,
                      const float* w_init, const float* b_init, int64_t n_atoms, int64_t n_bonds,
                      int32_t fa, int32_t fb, int32_t hidden, int32_t act, float* h0, float* z0,
                      void* workspace, size_t workspace_bytes, void* stream);

/* One directed-bond update, GNN.py:91-102 + DMPNNConv.forward :131-141:
 *   m[e] = sum_{k in in(src e)} h[k] - h[e^1];  z = W m + b + skip*h0;  h' = dropout(act(z)).
 * m_out [E,H] is required (it doubles as the saved message for backward); z_out may be NULL.
 * skip == NULL means weight 1.  Dropout is applied iff training && dropout_p > 0. */
int cgr_bond_update_fwd(const float* h_in, const float* h0, const int32_t* in_ptr,
                        const int32_t* in_idx, const int32_t* src, const float* w, const float* b,
                        const float* skip, int32_t act, float dropout_p, uint64_t seed,
                        uint32_t layer, int32_t training, float* h_out, float* m_out, float* z_out,
                        int64_t n_bonds, int64_t n_atoms, int32_t hidden, void* stream);

/* Stand-alone DMPNNConv.forward, GNN.py:131-141: a_out[v] = sum_{e in in(v)} h[e] (a_message) and
 * y_out[e] = W (a[src e] - h[e^1]) + b.  m_ws is [E,H] scratch. */
int cgr_conv_fwd(const float* h, const int32_t* in_ptr, const int32_t* in_idx, const int32_t* src,
                 const float* w, const float* b, float* a_out, float* y_out, float* m_ws, int64_t n_bonds,
                 int64_t n_atoms, int32_t hidden, void* stream);

/* Readout, GNN.py:105-110: s[v] = sum_{e in in(v)} h[e]; hv = act(W_o [x || s] + b_o);
 * out[b] = w_f . sum_{v in b} hv[v] + b_f.  s_out, hv_out, pooled_out are required scratch /
 * saved tensors; zv_out may be NULL. */
int cgr_readout_fwd(const float* h, const float* x, const int32_t* in_ptr, const int32_t* in_idx,
                    const int32_t* atom_ptr, const float* w_e2n, const float* b_e2n,
                    const float* w_ffn, const float* b_ffn, int32_t act, float* out, float* s_out,
                    float* hv_out, float* zv_out, float* pooled_out, int64_t n_atoms,
                    int64_t n_bonds, int64_t n_rxn, int32_t fa, int32_t hidden, void* stream);

/* ------------------------------------------------------------------------------------------
 * (5) Stage-level backward entry points: the explicit mirror of autograd for each stage above (exact-fp32 kernels,
 * fixed-order reductions: deterministic).  Every gradient buffer is written, not accumulated, except dh0_acc.
 * workspace >= cgr_stage_bwd_workspace(N, E, fa, fb, H) bytes.
 * ---------------------------------------------------------------------------------------- */
size_t cgr_stage_bwd_workspace(int64_t n_atoms, int64_t n_bonds, int32_t fa, int32_t fb, int32_t hidden);

/* Backward of cgr_readout_fwd (GNN.py:105-110): grad_out [B] and the saved s / hv / zv (NULL for relu) / pooled ->
 * gw_e2n [H, fa+H], gb_e2n [H], gw_ffn [H], gb_ffn [1] and dh [E, H], the gradient w.r.t. the bond states the readout
 * consumed (dh[e] = ds[dst e]). */
int cgr_readout_bwd(const float* grad_out, const float* x, const int32_t* in_ptr, const int32_t* in_idx,
                    const int32_t* atom_ptr, const int32_t* dst, const float* w_e2n, const float* w_ffn, int32_t act,
                    const float* s, const float* hv, const float* zv, const float* pooled, float* gw_e2n,
                    float* gb_e2n, float* gw_ffn, float* gb_ffn, float* dh, int64_t n_atoms, int64_t n_bonds,
                    int64_t n_rxn, int32_t fa, int32_t hidden, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of cgr_bond_update_fwd (GNN.py:91-102, 131-141): dh_out [E, H] (gradient w.r.t. the layer's output) with the
 * saved h_out, z (NULL for relu), m and h0 -> gw [H, H], gb [H], gskip (scalar, may be NULL), dh_in [E, H] (gradient
 * w.r.t. the layer's input: dh[k] = sum_{j in in(dst k)} dm[j^1] - dm[k^1], dm = dz W) and dh0_acc [E, H] (+)= skip * dz
 * (dh0_first != 0: stored instead of accumulated).  Dropout as in the forward (same seed / layer). */
int cgr_bond_update_bwd(const float* dh_out, const float* h_out, const float* z, const float* m, const float* h0,
                        const int32_t* in_ptr, const int32_t* in_idx, const int32_t* dst, const float* w,
                        const float* skip, int32_t act, float dropout_p, uint64_t seed, uint32_t layer,
                        int32_t training, float* gw, float* gb, float* gskip, float* dh_in, float* dh0_acc,
                        int32_t dh0_first, int64_t n_bonds, int64_t n_atoms, int32_t hidden, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Backward of cgr_edge_init_fwd (GNN.py:85-86): dh0 [E, H] (total gradient w.r.t. h_0: layer 0's input gradient plus the
 * accumulated skip contributions) with the saved h0 and z0 (NULL for relu) -> gw_init [H, fa+fb], gb_init [H]. */
int cgr_edge_init_bwd(const float* dh0, const float* h0, const float* z0, const float* x, const float* edge_attr,
                      const int32_t* in_ptr, const int32_t* in_idx, int32_t act, float* gw_init, float* gb_init,
                      int64_t n_atoms, int64_t n_bonds, int32_t fa, int32_t fb, int32_t hidden, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-network entry points (what the torch custom ops call).
 * ---------------------------------------------------------------------------------------- */
size_t cgr_forward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int32_t training,
                             int32_t engine);
/* out [B].  `saved` must be non-NULL when training != 0 (activations for backward).
 * Engine CGR_ENGINE_TC picks its path from the arguments:
 *   - g->tile_info set, saved == NULL           : fused tile kernels, inference: the atom projection, then ONE cluster
 *                                                 kernel (edge initialisation, every bond layer, readout, pool, FFN);
 *   - g->tile_info set, saved->tc_blob != NULL  : the same kernels keeping every layer's FP16 operands in tc_blob
 *                                                 (ReLU networks; cgr_tc_saved_bytes > 0) -- cgr_gnn_backward with the
 *                                                 same `saved` then runs the fused tile-local backward; when
 *                                                 p->tc_weights is NULL the prepared weights are stored in tc_blob too
 *                                                 and the backward reads them from there (pass the same p);
 *   - otherwise                                 : layer-wise kernels with tensor-core GEMMs (any activation, any
 *                                                 reaction size), activations in the fp32 buffers of `saved`. */
int cgr_gnn_forward(const cgr_params_t* p, const cgr_graph_t* g, float* out, cgr_saved_t* saved,
                    int32_t training, uint64_t seed, int32_t engine, void* workspace,
                    size_t workspace_bytes, void* stream);

/* The same inference forward for SEVERAL independent batches in one call (a screening loop -- test.py:103-113, the
 * CLI -- is a stream of them): `graphs[0..n_graphs)` (1..24 per call, every one with a tile plan, tc_status and,
 * optionally, prepared x_hi / x_lo), energies of batch i to outs[i] [graphs[i].n_rxn].  TWO launches per call instead
 * of two per batch: one atom projection over every batch's atom tiles and one fused cluster kernel over every batch's
 * tile groups (per-batch operands and index arrays travel as kernel parameters; the weights are shared), so a group
 * of 64-reaction batches fills the 148 SMs wave after wave.  Same kernels and arithmetic as cgr_gnn_forward with
 * p->tc_throughput = 1: bit-identical energies.  tcgen05 engine, inference only; p->tc_weights may be NULL
 * (prepared per call).  workspace >= cgr_forward_group_workspace() (0: a batch has no tile plan / n_graphs > 24). */
size_t cgr_forward_group_workspace(const cgr_params_t* p, const cgr_graph_t* graphs, int32_t n_graphs);
int cgr_gnn_forward_group(const cgr_params_t* p, const cgr_graph_t* graphs, int32_t n_graphs, float* const* outs,
                          void* workspace, size_t workspace_bytes, void* stream);

size_t cgr_backward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int32_t engine);
/* (5) explicit backward of (2)-(4): grad_out [B] -> every parameter gradient. */
int cgr_gnn_backward(const cgr_params_t* p, const cgr_graph_t* g, const cgr_saved_t* saved,
                     const float* grad_out, cgr_grads_t* grads, uint64_t seed, int32_t engine,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * tcgen05 engine helpers.
 * cgr_tc_plan_build packs consecutive whole reactions into 128-bond row tiles: tile_info[t] =
 * {first bond, #bonds, first atom, #atoms, first reaction, #reactions, 0, 0}; status[0] = number of
 * tiles, status[1] = 1 iff every reaction fits a tile (<= 128 bonds).  tile_info needs room for
 * n_rxn tiles.  cgr_tc_plan_check (after reading status[0]) clears status[1] if a bond leaves its tile.
 * cgr_tc_prepare_weights converts the parameter tree into the FP16 (hi, lo) operand layout once;
 * pass the buffer as cgr_params_t.tc_weights to skip the per-call conversion.
 * ---------------------------------------------------------------------------------------- */
int cgr_tc_plan_build(const int32_t* in_ptr, const int32_t* atom_ptr, int64_t n_rxn, int32_t* tile_info,
                      int32_t* status, void* stream);
int cgr_tc_plan_check(const int32_t* tile_info, int64_t n_tiles, const int32_t* src, const int32_t* dst,
                      int32_t* status, void* stream);
/* Test entry of the training GEMM (tc_gemm2): c[m,n] = sum_k A(m,k) B(n,k) on fp32 inputs through the scaled FP16x3
 * split; a_mn != 0: `a` is stored [k, m] (MN-major, as in weight gradients), else [m, k]; same for b. */
size_t cgr_tc_gemm_test_workspace(int64_t m, int64_t n, int64_t k);
int cgr_tc_gemm_test(const float* a, const float* b, int64_t m, int64_t n, int64_t k, int32_t a_mn, int32_t b_mn,
                     float* c, void* workspace, size_t workspace_bytes, void* stream);

/* Batch preparation for the tcgen05 engine: data.x [N, fa] fp32 -> FP16 (hi, lo) rows of stride
 * cgr_tc_features_ld(fa) halfs (part of collation, like the CSR arrays; status[0] gets the overflow flag). */
int64_t cgr_tc_features_ld(int32_t fa);
int cgr_tc_split_features(const float* x, int64_t n_atoms, int32_t fa, void* x_hi, void* x_lo,
                          int32_t* status, void* stream);
/* Debug: device buffer [n_cta][8] of int64 receiving clock64 stamps of the bond-layer kernel phases (NULL = off). */
int cgr_tc_debug_buffer(void* device_buffer);
size_t cgr_tc_weights_bytes(const cgr_params_t* p);
int cgr_tc_prepare_weights(const cgr_params_t* p, void* buffer, size_t buffer_bytes, void* stream);
/* Test entry: out[M,N] = x[M,K] w[N,K]^T + bias (bias may be NULL) on the TMA + tcgen05 FP16x3 pipeline. */
size_t cgr_tc_linear_workspace(int64_t m, int64_t n, int64_t k);
int cgr_tc_linear(const float* x, int64_t m, int64_t k, const float* w, int64_t n, const float* bias,
                  float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * End-to-end inference on HOST buffers (what test.py:111-113 / the CLI :71-76 do per batch):
 * stages x / edge_attr / edge_index to the device, builds the index arrays and the tile plan, runs the
 * tcgen05 forward and copies the energies back to `host_out` [B]; returns after the stream is idle.
 * host_ptr [B+1] (Batch.ptr) may be NULL when host_batch [N] is given (or both NULL for one graph).
 * `dev_ws` / `host_ws` (pinned) sizes from cgr_infer_host_workspace.  p->tc_weights must be prepared.
 * Returns CGR_ERR_UNSUPPORTED (-3) when a reaction does not fit a 128-bond tile (use the generic path).
 * ---------------------------------------------------------------------------------------- */
int cgr_infer_host_workspace(const cgr_params_t* p, int64_t n_atoms, int64_t n_bonds, int64_t n_rxn,
                             size_t* dev_bytes, size_t* host_bytes);
int cgr_gnn_infer_host(const cgr_params_t* p, const float* host_x, const float* host_edge_attr,
                       const int64_t* host_edge_index, const int64_t* host_ptr, const int64_t* host_batch,
                       int64_t n_atoms, int64_t n_bonds, int64_t n_rxn, float* host_out, void* dev_ws,
                       size_t dev_bytes, void* host_ws, size_t host_bytes, void* stream);

/* Asynchronous flavour for pipelining several batches over streams: enqueues everything on `stream` and
 * returns; after the stream is synchronised, cgr_infer_host_check decodes the validity flags in host_ws. */
int cgr_gnn_infer_host_async(const cgr_params_t* p, const float* host_x, const float* host_edge_attr,
                             const int64_t* host_edge_index, const int64_t* host_ptr, const int64_t* host_batch,
                             int64_t n_atoms, int64_t n_bonds, int64_t n_rxn, float* host_out, void* dev_ws,
                             size_t dev_bytes, void* host_ws, size_t host_bytes, void* stream);
int cgr_infer_host_check(const cgr_params_t* p, int64_t n_atoms, int64_t n_bonds, int64_t n_rxn,
                         const void* host_ws);

/* Several collated host batches in ONE submission (dynamic batching of a screening stream, CLI :71-76 looped over
 * a DataLoader): every batch is staged straight from its own host buffers into one device-side super-batch (atom
 * ids shifted on the device), so launches and copies are amortised over all of them.  Energies land in `host_out`
 * in batch order ([sum of n_rxn]); per-reaction results equal separate submissions up to fp32 rounding of the final
 * column sum (its order follows the N-slice width chosen for the super-batch).  Workspaces are
 * sized by cgr_infer_host_workspace on the TOTAL atom / bond / reaction counts, which cgr_infer_host_check takes too. */
typedef struct {
  const float* x;              /* [n_atoms, fa] */
  const float* edge_attr;      /* [n_bonds, fb] */
  const int64_t* edge_index;   /* [2, n_bonds] batch-local atom ids */
  const int64_t* ptr;          /* [n_rxn + 1] or NULL */
  const int64_t* batch;        /* [n_atoms] or NULL */
  int64_t n_atoms, n_bonds, n_rxn;
} cgr_host_batch_t;
int cgr_gnn_infer_host_multi_async(const cgr_params_t* p, const cgr_host_batch_t* batches, int32_t n_batches,
                                   float* host_out, void* dev_ws, size_t dev_bytes, void* host_ws,
                                   size_t host_bytes, void* stream);

/* Host-side tile plan for callers that know the per-reaction offsets (HOST arrays atom_ptr / edge_ptr [n_rxn + 1]):
 * the greedy packing of cgr_tc_plan_build without device work or a synchronisation.  tile_info: HOST [n_rxn][8] ints,
 * to be uploaded as cgr_graph_t.tile_info.  Returns CGR_ERR_UNSUPPORTED (-3) when a reaction exceeds a 128-row tile. */
int cgr_tc_plan_host(const int64_t* atom_ptr, const int64_t* edge_ptr, int64_t n_rxn, int32_t* tile_info,
                     int64_t* n_tiles);

/* Bytes of cgr_saved_t.tc_blob for the fused tile-local training path of the tcgen05 engine, or 0 when the
 * configuration cannot use it (needs ReLU, a tile plan in `g`, hidden % 4 == 0 and <= 1024): then leave tc_blob
 * NULL and provide the layer-wise buffers. */
size_t cgr_tc_saved_bytes(const cgr_params_t* p, const cgr_graph_t* g);

/* ------------------------------------------------------------------------------------------
 * Device-resident reaction store.  The featurised data set is kept packed in device memory
 * -- x_all [N_all, fa], ea_all [E_all, fb], ei_all [2, E_all] reaction-LOCAL atom ids (int32), node_ptr / edge_ptr
 * [R + 1] (int64), y_all [R] -- and one call assembles the batch of the reactions `sel` [n_sel] (any order, repeats
 * allowed): x, edge_attr, edge_index [2, e_out] (int64, batch-global ids), batch [n_out], y [n_sel].  The caller
 * supplies the output offsets out_node_ptr / out_edge_ptr [n_sel + 1] (exclusive scans of the selected sizes; they
 * are the batch's `ptr` / `edge_ptr`), all device pointers.  Replaces ChemDataset.__getitem__ + molgraph2data
 * (data/ChemDataset.py:69-94, float32 concatenation of CGR and MACE features) and PyG's collate
 * (training/trainer.py:105-118) for a data set that fits in HBM; results are bit-identical to the host collate.
 * ---------------------------------------------------------------------------------------- */
int cgr_store_gather(const float* x_all, const float* ea_all, const int32_t* ei_all, const int64_t* node_ptr,
                     const int64_t* edge_ptr, const float* y_all, int64_t e_all, const int64_t* sel,
                     const int64_t* out_node_ptr, const int64_t* out_edge_ptr, int64_t n_sel, int32_t fa, int32_t fb,
                     int64_t e_out, float* x, float* edge_attr, int64_t* edge_index, int64_t* batch, float* y,
                     void* stream);

/* Inference over a resident store in ONE call (the screening loop of test.py:103-113 / the CLI over a data set that
 * lives in HBM): for every consecutive `batch_size` ids of `order` (HOST, [n_total]) it assembles the batch on the
 * device, builds the index arrays and runs the tcgen05 forward, writing the energies to `out` (DEVICE, [n_total], in
 * `order` order).  Batches are pipelined over `n_slots` streams (`streams`: HOST array of cudaStream_t), each with its
 * own slice of `dev_ws` (n_slots * dev_bytes_per_slot) and of the pinned `host_ws` (n_slots * host_bytes_per_slot), both
 * sized by cgr_store_infer_workspace for the largest batch.  Returns after every stream is idle; -3 when a reaction
 * does not fit a 128-row tile.  p->tc_weights must be prepared.
 * Inside the call the loop is free in everything the caller cannot observe: consecutive small batches are assembled as
 * super-batches of <= 1024 reactions, the reactions of a batch are assembled in best-fit order for the 128-row tiles
 * (energies scattered back to `order` positions), and the atom features go from the store rows straight to the FP16
 * (hi, lo) operands of the atom projection (no fp32 copy of x).  A reaction's energy depends on neither its batch nor its
 * position, so `out` equals the forward of the caller-ordered batches in the same kernel configuration bit for bit
 * (and any other configuration within the parity tolerance). */
typedef struct {
  const float* x_all;            /* device */
  const float* ea_all;
  const int32_t* ei_all;
  const int64_t* node_ptr;
  const int64_t* edge_ptr;
  const int64_t* node_ptr_host;  /* host copies of the two offset arrays */
  const int64_t* edge_ptr_host;
  int64_t n_rxn, e_all;
  int32_t fa, fb;
} cgr_store_t;
/* The order in which cgr_store_infer assembles the reactions `ids` [n] (HOST) of one batch: bucketed best-fit decreasing
 * on the directed-bond counts under the atoms <= 128 constraint, so that the greedy tile plan over consecutive reactions
 * (cgr_tc_plan_host) closes every 128-row tile nearly full (T1x-shaped batches: ~97 % instead of ~87 %).  perm [n]
 * (HOST, out): perm[i] = position in `ids` of the reaction assembled at position i.  Pure host code.  Returns -3 (and the
 * identity order) when a reaction does not fit a tile.  No reference counterpart: PyG collates in loader order
 * (training/trainer.py:105-118), and the energies do not depend on it (GNN.py:110 pools per reaction). */
int cgr_store_pack_order(const int64_t* node_ptr_host, const int64_t* edge_ptr_host, int64_t n_rxn_store,
                         const int64_t* ids, int64_t n, int32_t* perm);
int cgr_store_infer_workspace(const cgr_params_t* p, const cgr_store_t* store, const int64_t* order, int64_t n_total,
                              int64_t batch_size, size_t* dev_bytes_per_slot, size_t* host_bytes_per_slot);
int cgr_store_infer(const cgr_params_t* p, const cgr_store_t* store, const int64_t* order, int64_t n_total,
                    int64_t batch_size, float* out, void* dev_ws, size_t dev_bytes_per_slot, void* host_ws,
                    size_t host_bytes_per_slot, int32_t n_slots, void* const* streams);

/* ------------------------------------------------------------------------------------------
 * CGR featurisation after the chemistry toolkit (reference utils/graph_features.py:4-63 atom /
 * bond features, :177-195 reactant || (product - reactant) layout).  The caller parses SMILES (RDKit in the reference) and
 * hands over compact attribute codes, product side aligned to the reactant atom order (graph_features.py:83-103):
 *   atom_r / atom_p  [N, 6] int16: atomic number, total degree, formal charge, total #Hs, hybridisation code, aromatic;
 *   mass_r / mass_p  [N] double (atom.GetMass(); the reference multiplies by 0.01 in Python floats);
 *   bond_r / bond_p  [E, 3] int8 per DIRECTED bond of the reactant/product union, in the reference's edge order:
 *                    type (-1 absent on this side, 0 single, 1 double, 2 triple, 3 aromatic, other = another type),
 *                    conjugated, in ring.
 * `host_tables` holds the one-hot choice lists (table-driven: the reference's lists are the defaults in
 * cgr_mpnn_3d_b200/featurize.py; values not listed light the extra last slot, graph_features.py:66-80).
 * Writes x[v, 0:78] (row stride ldx >= 78 floats, so a MACE block can follow in the same tensor) and edge_attr [E, 14].
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int16_t symbol_z[11];        /* H C N O F Si P S Cl Br I as atomic numbers (graph_features.py:16-18) */
  int16_t degrees[6];
  int16_t charges[5];
  int16_t num_hs[5];
  int16_t hybridizations[5];   /* the parser's codes for SP, SP2, SP3, SP3D, SP3D2 */
} cgr_feature_tables_t;
int cgr_featurize_cgr(const cgr_feature_tables_t* host_tables, const int16_t* atom_r, const int16_t* atom_p,
                      const double* mass_r, const double* mass_p, int64_t n_atoms, const int8_t* bond_r,
                      const int8_t* bond_p, int64_t n_bonds, float* x, int64_t ldx, float* edge_attr, void* stream);

/* Loss adjacent to the path (train.py:120, trainer.py:142): L = sum_b (pred-y)^2, and dL/dpred. */
int cgr_mse_sum_fwd_bwd(const float* pred, const float* y, int64_t n_rxn, float* loss,
                        float* grad_pred, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused optimizer step: what `optimizer.step()` does at training/trainer.py:144 with the
 * optimizer of train.py:117-119 -- torch.optim.Adam(lr, weight_decay (L2 added to the gradient), amsgrad=True) --
 * for ALL parameter tensors in one launch.  `tensors` is a HOST array; every pointer in it is a device pointer to
 * contiguous fp32.  `step` counts from 1 (bias corrections 1 - beta^step, computed in double like the reference).
 * `grad_scale` multiplies every gradient first (1.0, or 1/world for a mean over replicas).
 * max_exp_avg_sq may be NULL when amsgrad == 0.  The learning-rate schedule (ExponentialLR, train.py:121) stays
 * with the caller: pass the current lr.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  float* max_exp_avg_sq;
  int64_t numel;
} cgr_adam_tensor_t;
int cgr_adam_step(const cgr_adam_tensor_t* tensors, int32_t n_tensors, double lr, double beta1, double beta2,
                  double eps, double weight_decay, int64_t step, int32_t amsgrad, float grad_scale, void* stream);

/* Data-parallel optimizer step WITHOUT a collective library: the SUM of the replicas' gradients (the loss is
 * MSELoss(reduction="sum"), train.py:120, so replicas sum) and the Adam update of train.py:117-119 in ONE kernel.
 * `peer_arenas` / `peer_flags`: HOST arrays [world] of DEVICE pointers -- every rank's flat fp32 gradient arena of this
 * step and flag pad (int[64], zero-initialised once), rank order, the caller's own included; the peers' ones are mapped
 * with CUDA IPC and reachable after cgr_enable_peer_access.  In `tensors`, `grad` holds the tensor's OFFSET inside the
 * arenas in floats (cast to a pointer), identical on every rank.  `sync_step` must grow by one per call on every rank
 * (1, 2, ...): the kernel announces it in all flag pads, waits for all peers, then reads each gradient element from all
 * arenas in rank order (bit-identical sums on every replica) and updates its own parameters.  Callers alternate two
 * arenas per rank so a fast rank never overwrites gradients a slow peer still reads.
 * `peer_reduced` NULL: one-shot (every rank reads all W arenas).  Non-NULL (HOST array [world] of DEVICE pointers to
 * every rank's `arena_floats`-long scratch buffer, mapped like the arenas): two-shot -- each rank reduces one slice, all
 * ranks gather the slices: 2 (W-1)/W instead of W-1 gradient sizes over NVLink per GPU, for larger replica counts. */
int cgr_enable_peer_access(int32_t peer_device);
/* CUDA IPC helpers for the arenas: export = 64-byte handle of the allocation containing `dev_ptr` + the pointer's byte
 * offset inside it; open = map a peer's allocation for the calling thread's current device and return the address of the
 * same bytes (the mapping stays open for the life of the process). */
int cgr_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out);
int cgr_ipc_open(const void* handle, int64_t offset, void** ptr_out);
int cgr_peer_allreduce_adam(const cgr_adam_tensor_t* tensors, int32_t n_tensors, const float* const* peer_arenas,
                            int* const* peer_flags, float* const* peer_reduced, int64_t arena_floats, int32_t world,
                            int32_t rank, int32_t sync_step, double lr, double beta1, double beta2, double eps,
                            double weight_decay, int64_t step, int32_t amsgrad, float grad_scale, void* stream);

/* Measurement hooks used by bench.py: number of kernels this library has launched so far, and
 * optional CUDA-event timing of each named stage (events are recorded on the launching stream). */
long long cgr_launch_count(void);
int cgr_profile_enable(int enable);            /* also clears previously recorded ranges */
int cgr_profile_count(void);
int cgr_profile_get(int i, char* name_out, int name_cap, float* ms_out);

/* Debug / test helper: the dropout keep-mask (uint8 [E,H]) the kernels use for `layer`. */
int cgr_dropout_mask(uint64_t seed, uint32_t layer, float dropout_p, int64_t n_bonds,
                     int32_t hidden, uint8_t* mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGR_B200_H_ */
