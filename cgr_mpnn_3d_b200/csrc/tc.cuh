// tcgen05 (5th-gen tensor core) engine: FP16x3-split GEMMs with TMEM accumulators, TMA-fed weights,
// gathers fused around the MMA.  Declarations shared with api.cu.
#pragma once
#include "../../include/cgr_b200.h"
#include "common.cuh"

size_t tc_forward_workspace(const cgr_params_t* p, const cgr_graph_t* g, int training);
int tc_gnn_forward(const cgr_params_t* p, const cgr_graph_t* g, float* out, cgr_saved_t* saved, int training,
                   uint64_t seed, void* workspace, size_t workspace_bytes, cudaStream_t st);
