// tcgen05 (5th-gen tensor core) engine: FP16x3-split GEMMs with TMEM accumulators, TMA-fed weights,
// gathers fused around the MMA.  Declarations shared with api.cu.
#pragma once
#include "../../include/cgr_b200.h"
#include "common.cuh"

size_t tc_forward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int training);
int tc_gnn_forward(const cgr_params_t* p, const cgr_graph_t* g, float* out, cgr_saved_t* saved, int training,
                   uint64_t seed, void* workspace, size_t workspace_bytes, cudaStream_t st);

size_t tc_weights_bytes(const cgr_params_t* p);
int tc_prepare_weights(const cgr_params_t* p, void* wbuf, size_t wbuf_bytes, cudaStream_t st);
int tc_plan_build(const int32_t* in_ptr, const int32_t* atom_ptr, const int32_t* src, const int32_t* dst,
                  int64_t n_rxn, int32_t* tile_info, int32_t* status, cudaStream_t st);
int tc_plan_check(const int32_t* tile_info, int64_t n_tiles, const int32_t* src, const int32_t* dst, int32_t* status,
                  cudaStream_t st);
size_t tc_linear_workspace(int64_t M, int64_t N, int64_t K);
int tc_linear(const float* x, int64_t M, int64_t K, int64_t ldx, const float* wgt, int64_t N, int64_t ldw,
              const float* bias, float* out, void* workspace, size_t workspace_bytes, cudaStream_t st);
void tc_set_debug_buffer(long long* p);
int tc_split_features(const float* x, int64_t n, int fa, void* x_hi, void* x_lo, int* status, cudaStream_t st);
