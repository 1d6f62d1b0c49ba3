// tcgen05 (5th-gen tensor core) engine: FP16x3-split GEMMs with TMEM accumulators, TMA-fed weights,
// gathers fused around the MMA.  Declarations shared with api.cu.
#pragma once
#include "../../include/cgr_b200.h"
#include "common.cuh"
#include "simt.cuh"

size_t tc_forward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int training);
size_t tc_forward_group_workspace(const cgr_params_t* p, const cgr_graph_t* gs, int n);
int tc_gnn_forward_group(const cgr_params_t* p, const cgr_graph_t* gs, int n, float* const* outs, void* workspace,
                         size_t workspace_bytes, cudaStream_t st);
int tc_gnn_forward(const cgr_params_t* p, const cgr_graph_t* g, float* out, cgr_saved_t* saved, int training,
                   uint64_t seed, void* workspace, size_t workspace_bytes, cudaStream_t st);

// fused tile-local training path (ReLU, tileable batches): activations kept in cgr_saved_t.tc_blob
bool tc_fused_training_ok(const cgr_params_t* p, const cgr_graph_t* g);
size_t tc_saved_bytes(const cgr_params_t* p, const cgr_graph_t* g);
size_t tc_backward_workspace(const cgr_params_t* p, const cgr_graph_t* g);
int tc_gnn_backward(const cgr_params_t* p, const cgr_graph_t* g, const cgr_saved_t* saved, const float* dout,
                    const cgr_grads_t* grads, void* workspace, size_t workspace_bytes, cudaStream_t st);

size_t tc_weights_bytes(const cgr_params_t* p);
int tc_prepare_weights(const cgr_params_t* p, void* wbuf, size_t wbuf_bytes, cudaStream_t st);
int tc_plan_build(const int32_t* in_ptr, const int32_t* atom_ptr, const int32_t* src, const int32_t* dst,
                  int64_t n_rxn, int32_t* tile_info, int32_t* status, cudaStream_t st);
int tc_plan_host(const int64_t* atom_ptr, const int64_t* edge_ptr, int64_t n_rxn, int32_t* tile_info, int64_t* n_tiles);
int tc_plan_check(const int32_t* tile_info, int64_t n_tiles, const int32_t* src, const int32_t* dst, int32_t* status,
                  cudaStream_t st);
size_t tc_linear_workspace(int64_t M, int64_t N, int64_t K);
int tc_linear(const float* x, int64_t M, int64_t K, int64_t ldx, const float* wgt, int64_t N, int64_t ldw,
              const float* bias, float* out, void* workspace, size_t workspace_bytes, cudaStream_t st);
void tc_set_debug_buffer(long long* p);
// tc_status[0] handling of the layer-wise path: clear the per-forward overflow bit / poison the energies with NaN when set
int tc_flag_begin(int* flag, cudaStream_t st);
int tc_poison_outputs(float* out, int64_t n, const int* flag, cudaStream_t st);
int tc_split_features(const float* x, int64_t n, int fa, void* x_hi, void* x_lo, int* status, cudaStream_t st);

// ---- training path: tensor-core GEMMs on FP16 (hi, lo) operands of either major ----
struct TcOperand {
  const __half* hi;
  const __half* lo;
  int64_t ld;              // row stride in halfs (multiple of 8)
  const float* unscale;    // device scalar 1/scale or null
  bool mn_major;           // false: [MN rows, K cols];  true: [K rows, MN cols]
};
int tc_split(const float* in, int64_t ld, int64_t rows, int cols, bool scaled, __half* hi, __half* lo, int64_t ldo,
             unsigned int* amax_slot, float* unscale_slot, int* overflow, cudaStream_t st);
int tc_splitk_choose(int64_t M, int64_t N, int64_t K);
int tc_train_gemm(const TcOperand& A, const TcOperand& B, int64_t M, int64_t N, int64_t K, float* C, int64_t ldc,
                  const GemmEpilogue& epi, int split_k, float* partial, cudaStream_t st);
int tc_splitk_batched(int n, int64_t M, int64_t N, int64_t K);
int tc_train_gemm_batched_mn(const TcOperand* A, const TcOperand* B, float* const* C, const int64_t* ldc, int n,
                             int64_t M, int64_t N, int64_t K, int split_k, float* partial, const char* name,
                             cudaStream_t st);
TcOperand tc_weight_operand(const cgr_params_t* p, const void* wbuf, int mat, int64_t row0, bool mn_major);
size_t tc_gemm2_test_workspace(int64_t M, int64_t N, int64_t K);
int tc_gemm2_test(const float* A, const float* B, int64_t M, int64_t N, int64_t K, int a_mn, int b_mn, float* C,
                  void* workspace, size_t workspace_bytes, cudaStream_t st);
