// Shared device/host helpers for libcgr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define CGR_OK 0
#define CGR_ERR_ARG (-1)
#define CGR_ERR_WORKSPACE (-2)
#define CGR_ERR_UNSUPPORTED (-3)

// thread-local last-error text, read through cgr_last_error_string()
void cgr_set_error(const char* fmt, ...);

#define CGR_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      cgr_set_error(__VA_ARGS__);                \
      return CGR_ERR_ARG;                        \
    }                                            \
  } while (0)

#define CGR_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) {                                                  \
      cgr_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                    cudaGetErrorString(_e));                                  \
      return (int)_e;                                                         \
    }                                                                         \
  } while (0)

#define CGR_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) {                                                  \
      cgr_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,  \
                    cudaGetErrorString(_e));                                  \
      return (int)_e;                                                         \
    }                                                                         \
  } while (0)

// launch accounting + optional per-kernel event timing (cgr_profile_* in the C ABI)
void cgr_note_launch(const char* name, cudaStream_t st, int n);
struct CgrRange {   // RAII: brackets the kernel launches of one named stage with CUDA events when profiling
  const char* name;
  cudaStream_t st;
  bool active;
  CgrRange(const char* name, cudaStream_t st);
  ~CgrRange();
};

static inline int64_t cgr_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t cgr_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// activation ids shared with the Python side (reference train.py:284-292 maps the CLI
// choice to F.relu / F.silu / F.gelu; F.gelu default is the exact erf form)
enum { CGR_ACT_RELU = 0, CGR_ACT_SILU = 1, CGR_ACT_GELU = 2, CGR_ACT_IDENTITY = 3 };

__device__ __forceinline__ float cgr_act(float z, int act) {
  switch (act) {
    case CGR_ACT_RELU: return z > 0.f ? z : 0.f;
    case CGR_ACT_SILU: return z / (1.f + expf(-z));
    case CGR_ACT_GELU: return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
    default: return z;
  }
}

// derivative of the activation w.r.t. its pre-activation z (h = act(z) given for relu)
__device__ __forceinline__ float cgr_act_grad(float z, float h, int act) {
  switch (act) {
    case CGR_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case CGR_ACT_SILU: {
      float s = 1.f / (1.f + expf(-z));
      return s * (1.f + z * (1.f - s));
    }
    case CGR_ACT_GELU: {
      float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
      float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    default: return 1.f;
  }
}

// ---- counter-based RNG for dropout: Philox4x32-10 keyed by (seed), counter = (element/4, layer) ----
__device__ __forceinline__ uint4 cgr_philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// keep-decision for element `idx` of layer `layer`: uniform u in [0,1) compared with p
__device__ __forceinline__ bool cgr_dropout_keep(uint64_t seed, uint32_t layer, uint64_t idx, float p) {
  uint4 ctr = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), layer, 0x43475242u);
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  uint4 r = cgr_philox4x32(ctr, key);
  uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  float u = (float)(w >> 8) * (1.0f / 16777216.0f);
  return u >= p;
}
