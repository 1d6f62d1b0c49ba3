// Fused optimizer step: Adam with L2-in-gradient weight decay and amsgrad, as the reference
// configures it (train.py:117-121 -- torch.optim.Adam(lr, weight_decay, amsgrad=True)), over every parameter tensor
// in ONE launch.  HBM-bound: per element it reads param, grad, m, v, vmax and writes param, m, v, vmax (36 bytes).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cgr_b200.h"
#include "common.cuh"

namespace {

constexpr int ADAM_MAX_TENSORS = 48;
constexpr int ADAM_CHUNK = 4096;          // elements per block
constexpr int ADAM_THREADS = 256;

struct AdamTable {
  float* param[ADAM_MAX_TENSORS];
  const float* grad[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  float* vmax[ADAM_MAX_TENSORS];
  long long numel[ADAM_MAX_TENSORS];
  int first_block[ADAM_MAX_TENSORS + 1];  // blocks [first_block[t], first_block[t+1]) work on tensor t
  int n;
};

struct AdamScalars {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, neg_step_size, bc2_sqrt, grad_scale;
  int amsgrad;
};

// One element, the operation order of torch.optim.Adam's single-tensor path (torch/optim/adam.py, _single_tensor_adam):
//   g += wd * p;  m.lerp_(g, 1-b1);  v = v*b2 + (1-b2)*g*g;  vmax = max(vmax, v);
//   denom = sqrt(vmax) / sqrt(bias_correction2) + eps;  p += (-lr / bias_correction1) * (m / denom)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float& vm, const AdamScalars& s) {
  g *= s.grad_scale;
  if (s.weight_decay != 0.f) g = fmaf(s.weight_decay, p, g);
  m = fmaf(s.one_minus_beta1, g - m, m);
  v = fmaf(s.one_minus_beta2 * g, g, v * s.beta2);
  float d;
  if (s.amsgrad) {
    vm = fmaxf(vm, v);
    d = sqrtf(vm) / s.bc2_sqrt + s.eps;
  } else {
    d = sqrtf(v) / s.bc2_sqrt + s.eps;
  }
  p = fmaf(s.neg_step_size, m / d, p);
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const __grid_constant__ AdamTable tab,
                                                            const __grid_constant__ AdamScalars s) {
  int lo = 0, hi = tab.n;                      // tensor of this block: last t with first_block[t] <= blockIdx.x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.first_block[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const int t = lo;
  const long long base = (long long)((int)blockIdx.x - tab.first_block[t]) * ADAM_CHUNK;
  const long long n = tab.numel[t];
  float* __restrict__ P = tab.param[t];
  const float* __restrict__ G = tab.grad[t];
  float* __restrict__ M = tab.m[t];
  float* __restrict__ V = tab.v[t];
  float* __restrict__ VM = tab.vmax[t];
  const bool vec = (((uintptr_t)P | (uintptr_t)G | (uintptr_t)M | (uintptr_t)V | (uintptr_t)(s.amsgrad ? VM : P)) & 15) == 0;
  if (vec && base + ADAM_CHUNK <= n) {
#pragma unroll
    for (int it = 0; it < ADAM_CHUNK / (ADAM_THREADS * 4); ++it) {
      const long long i = base + (long long)(it * ADAM_THREADS + threadIdx.x) * 4;
      float4 p = *reinterpret_cast<float4*>(P + i);
      const float4 g = *reinterpret_cast<const float4*>(G + i);
      float4 m = *reinterpret_cast<float4*>(M + i);
      float4 v = *reinterpret_cast<float4*>(V + i);
      float4 vm = s.amsgrad ? *reinterpret_cast<float4*>(VM + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      adam_one(p.x, g.x, m.x, v.x, vm.x, s);
      adam_one(p.y, g.y, m.y, v.y, vm.y, s);
      adam_one(p.z, g.z, m.z, v.z, vm.z, s);
      adam_one(p.w, g.w, m.w, v.w, vm.w, s);
      *reinterpret_cast<float4*>(P + i) = p;
      *reinterpret_cast<float4*>(M + i) = m;
      *reinterpret_cast<float4*>(V + i) = v;
      if (s.amsgrad) *reinterpret_cast<float4*>(VM + i) = vm;
    }
  } else {
    const long long end = base + ADAM_CHUNK < n ? base + ADAM_CHUNK : n;
    for (long long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
      float p = P[i], m = M[i], v = V[i], vm = s.amsgrad ? VM[i] : 0.f;
      adam_one(p, G[i], m, v, vm, s);
      P[i] = p; M[i] = m; V[i] = v;
      if (s.amsgrad) VM[i] = vm;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Data-parallel step without NCCL: gradient SUM over the replicas and the Adam update in ONE kernel.  Every rank keeps its
// flat gradient arena in memory its peers can map (CUDA IPC); the kernel announces "my gradients of step s are complete"
// in every peer's flag pad, waits until all peers have announced the same, then reads each gradient element from all
// arenas over NVLink (fixed rank order: every replica computes bit-identical sums and stays in lock-step) and applies
// Adam to its own parameters.  Arenas alternate between two buffers, so a rank may start writing step s+1 while slower
// peers still read step s.
// ------------------------------------------------------------------------------------------------
constexpr int PEER_MAX = 16;
struct PeerTable {
  const float* arena[PEER_MAX];     // every rank's gradient arena of this step (own one included), rank order
  int* flags[PEER_MAX];             // every rank's flag pad (own one included): [0,16) gradients ready, [16,32) slice
                                    // reduced (two-shot), [32] block counter of the owner
  float* reduced[PEER_MAX];         // two-shot: every rank's buffer holding ITS reduced slice (arena index space)
  long long total, slice;           // two-shot: arena length and slice length per rank (floats, multiples of 4)
  int total_blocks;                 // two-shot: virtual blocks of the update phase (the launch is persistent)
  int world, rank, step;
  unsigned long long timeout_ns;    // how long to wait for a peer before failing the launch (CGR_PEER_TIMEOUT_S, default 600 s)
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float ld_peer(const float* p) {       // peer lines must not be served from a stale L1 line
  float v;
  asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(ADAM_THREADS) peer_adam_kernel(const __grid_constant__ AdamTable tab,
                                                                 const __grid_constant__ AdamScalars s,
                                                                 const __grid_constant__ PeerTable pt) {
  // 1. announce (the gradients were written by earlier kernels of this stream: complete and visible device-wide; the
  //    system-scope fence orders them before the flag for the peers), 2. wait for every peer
  if (blockIdx.x == 0 && threadIdx.x < pt.world) {
    __threadfence_system();
    volatile int* f = pt.flags[threadIdx.x] + pt.rank;
    *f = pt.step;
  }
  if (threadIdx.x < pt.world) {
    volatile int* mine = pt.flags[pt.rank] + threadIdx.x;
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (*mine < pt.step) {
      // a peer that stays away for the whole timeout (default 10 minutes: validation, checkpointing and stragglers are
      // fine) is taken for dead: fail the launch instead of hanging the GPU
      if ((++spins & 1023u) == 0 && global_ns() - t0 > pt.timeout_ns) __trap();
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncthreads();

  int lo = 0, hi = tab.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.first_block[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const int t = lo;
  const long long base = (long long)((int)blockIdx.x - tab.first_block[t]) * ADAM_CHUNK;
  const long long n = tab.numel[t];
  float* __restrict__ P = tab.param[t];
  const long long goff = (long long)(size_t)tab.grad[t];      // gradient position inside the arenas, in floats
  float* __restrict__ M = tab.m[t];
  float* __restrict__ V = tab.v[t];
  float* __restrict__ VM = tab.vmax[t];
  const bool vec = ((((uintptr_t)P | (uintptr_t)M | (uintptr_t)V | (uintptr_t)(s.amsgrad ? VM : P)) & 15) == 0) && (goff & 3) == 0;
  if (vec && base + ADAM_CHUNK <= n) {
#pragma unroll
    for (int it = 0; it < ADAM_CHUNK / (ADAM_THREADS * 4); ++it) {
      const long long i = base + (long long)(it * ADAM_THREADS + threadIdx.x) * 4;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < pt.world; ++r) {
        const float4 x = ld_peer4(pt.arena[r] + goff + i);
        g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
      }
      float4 p = *reinterpret_cast<float4*>(P + i);
      float4 m = *reinterpret_cast<float4*>(M + i);
      float4 v = *reinterpret_cast<float4*>(V + i);
      float4 vm = s.amsgrad ? *reinterpret_cast<float4*>(VM + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      adam_one(p.x, g.x, m.x, v.x, vm.x, s);
      adam_one(p.y, g.y, m.y, v.y, vm.y, s);
      adam_one(p.z, g.z, m.z, v.z, vm.z, s);
      adam_one(p.w, g.w, m.w, v.w, vm.w, s);
      *reinterpret_cast<float4*>(P + i) = p;
      *reinterpret_cast<float4*>(M + i) = m;
      *reinterpret_cast<float4*>(V + i) = v;
      if (s.amsgrad) *reinterpret_cast<float4*>(VM + i) = vm;
    }
  } else {
    const long long end = base + ADAM_CHUNK < n ? base + ADAM_CHUNK : n;
    for (long long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
      float g = 0.f;
      for (int r = 0; r < pt.world; ++r) g += ld_peer(pt.arena[r] + goff + i);
      float p = P[i], m = M[i], v = V[i], vm = s.amsgrad ? VM[i] : 0.f;
      adam_one(p, g, m, v, vm, s);
      P[i] = p; M[i] = m; V[i] = v;
      if (s.amsgrad) VM[i] = vm;
    }
  }
}


__device__ __forceinline__ void peer_wait(volatile int* flags, int world, int step, unsigned long long timeout_ns) {
  if ((int)threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (flags[threadIdx.x] < step) {
      if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) __trap();   // a peer died: fail the launch
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncthreads();
}

// Two-shot variant for larger replica counts: rank r first reduces only ITS slice of the gradient (reads that slice
// from all W arenas, rank order) into its `reduced` buffer, then every rank gathers the W reduced slices and updates its
// parameters -- 2 (W-1)/W gradient sizes over NVLink per GPU instead of W-1.  Persistent launch (all blocks co-resident:
// they wait for each other through the owner's block counter).
__global__ void __launch_bounds__(ADAM_THREADS) peer_adam2_kernel(const __grid_constant__ AdamTable tab,
                                                                  const __grid_constant__ AdamScalars s,
                                                                  const __grid_constant__ PeerTable pt) {
  int* myflags = pt.flags[pt.rank];
  // phase 0: gradients ready everywhere
  if (blockIdx.x == 0 && (int)threadIdx.x < pt.world) {
    __threadfence_system();
    volatile int* f = pt.flags[threadIdx.x] + pt.rank;
    *f = pt.step;
  }
  peer_wait(myflags, pt.world, pt.step, pt.timeout_ns);
  // phase 1: reduce my slice
  {
    const long long lo = (long long)pt.rank * pt.slice;
    long long hi = lo + pt.slice;
    if (hi > pt.total) hi = pt.total;
    float* out = pt.reduced[pt.rank];
    for (long long i = lo + ((long long)blockIdx.x * ADAM_THREADS + threadIdx.x) * 4; i < hi;
         i += (long long)gridDim.x * ADAM_THREADS * 4) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < pt.world; ++r) {
        const float4 x = ld_peer4(pt.arena[r] + i);
        g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
      }
      *reinterpret_cast<float4*>(out + i) = g;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(myflags + 32, 1);
    if (done == (int)gridDim.x - 1) {                 // every block of this rank has written its part of the slice
      myflags[32] = 0;
      __threadfence_system();
      for (int r = 0; r < pt.world; ++r) {
        volatile int* f = pt.flags[r] + 16 + pt.rank;
        *f = pt.step;
      }
    }
  }
  // phase 2: all slices reduced everywhere -> gather + Adam
  peer_wait(myflags + 16, pt.world, pt.step, pt.timeout_ns);
  for (int vb = blockIdx.x; vb < pt.total_blocks; vb += gridDim.x) {
    int lo = 0, hi = tab.n;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tab.first_block[mid] <= vb) lo = mid; else hi = mid;
    }
    const int t = lo;
    const long long base = (long long)(vb - tab.first_block[t]) * ADAM_CHUNK;
    const long long n = tab.numel[t];
    float* __restrict__ P = tab.param[t];
    const long long goff = (long long)(size_t)tab.grad[t];
    float* __restrict__ M = tab.m[t];
    float* __restrict__ V = tab.v[t];
    float* __restrict__ VM = tab.vmax[t];
    const bool vec = ((((uintptr_t)P | (uintptr_t)M | (uintptr_t)V | (uintptr_t)(s.amsgrad ? VM : P)) & 15) == 0) && (goff & 3) == 0;
    if (vec && base + ADAM_CHUNK <= n) {
#pragma unroll
      for (int it = 0; it < ADAM_CHUNK / (ADAM_THREADS * 4); ++it) {
        const long long i = base + (long long)(it * ADAM_THREADS + threadIdx.x) * 4;
        const long long a = goff + i;                        // slices are multiples of 4: one owner per float4
        const float4 g = ld_peer4(pt.reduced[(int)(a / pt.slice)] + a);
        float4 p = *reinterpret_cast<float4*>(P + i);
        float4 m = *reinterpret_cast<float4*>(M + i);
        float4 v = *reinterpret_cast<float4*>(V + i);
        float4 vm = s.amsgrad ? *reinterpret_cast<float4*>(VM + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        adam_one(p.x, g.x, m.x, v.x, vm.x, s);
        adam_one(p.y, g.y, m.y, v.y, vm.y, s);
        adam_one(p.z, g.z, m.z, v.z, vm.z, s);
        adam_one(p.w, g.w, m.w, v.w, vm.w, s);
        *reinterpret_cast<float4*>(P + i) = p;
        *reinterpret_cast<float4*>(M + i) = m;
        *reinterpret_cast<float4*>(V + i) = v;
        if (s.amsgrad) *reinterpret_cast<float4*>(VM + i) = vm;
      }
    } else {
      const long long end = base + ADAM_CHUNK < n ? base + ADAM_CHUNK : n;
      for (long long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
        const long long a = goff + i;
        const float g = ld_peer(pt.reduced[(int)(a / pt.slice)] + a);
        float p = P[i], m = M[i], v = V[i], vm = s.amsgrad ? VM[i] : 0.f;
        adam_one(p, g, m, v, vm, s);
        P[i] = p; M[i] = m; V[i] = v;
        if (s.amsgrad) VM[i] = vm;
      }
    }
  }
}

}  // namespace

extern "C" int cgr_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  CGR_CUDA(cudaGetDevice(&dev));
  if (dev == peer_device) return CGR_OK;
  int can = 0;
  CGR_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  CGR_CHECK_ARG(can, "device %d cannot map the memory of device %d (no P2P path)", dev, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return CGR_OK; }
  CGR_CUDA(e);
  return CGR_OK;
}

// CUDA IPC for the gradient arenas.  Export: the handle of the allocation that contains `dev_ptr` plus the pointer's
// offset inside it (torch tensors live inside larger cudaMalloc'ed segments).  Open: map a peer's allocation for the
// CURRENT device (peer access is enabled lazily by the driver) and return the address of the same bytes.
extern "C" int cgr_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  CGR_CHECK_ARG(dev_ptr && handle_out && offset_out, "cgr_ipc_export: null pointer");
  typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CGR_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q));
  CGR_CHECK_ARG(f && q == cudaDriverEntryPointSuccess, "cuMemGetAddressRange is not available from the driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = ((GetRangeFn)f)(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr);
  CGR_CHECK_ARG(r == CUDA_SUCCESS, "cuMemGetAddressRange failed with CUresult %d", (int)r);
  cudaIpcMemHandle_t h;
  CGR_CUDA(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  return CGR_OK;
}

extern "C" int cgr_ipc_open(const void* handle, int64_t offset, void** ptr_out) {
  CGR_CHECK_ARG(handle && ptr_out && offset >= 0, "cgr_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  CGR_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = (char*)base + offset;
  return CGR_OK;
}

extern "C" int cgr_peer_allreduce_adam(const cgr_adam_tensor_t* tensors, int32_t n_tensors, const float* const* peer_arenas,
                                       int* const* peer_flags, float* const* peer_reduced, int64_t arena_floats,
                                       int32_t world, int32_t rank, int32_t sync_step, double lr, double beta1,
                                       double beta2, double eps, double weight_decay, int64_t step, int32_t amsgrad,
                                       float grad_scale, void* stream) {
  CGR_CHECK_ARG(tensors && n_tensors > 0 && n_tensors <= ADAM_MAX_TENSORS, "cgr_peer_allreduce_adam: 1..%d tensors", ADAM_MAX_TENSORS);
  CGR_CHECK_ARG(peer_arenas && peer_flags && world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world,
                "cgr_peer_allreduce_adam: bad peer table");
  CGR_CHECK_ARG(step >= 1 && sync_step >= 1, "cgr_peer_allreduce_adam: steps count from 1");
  cudaStream_t st = (cudaStream_t)stream;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars s;
  s.beta1 = (float)beta1; s.beta2 = (float)beta2;
  s.one_minus_beta1 = (float)(1.0 - beta1); s.one_minus_beta2 = (float)(1.0 - beta2);
  s.eps = (float)eps; s.weight_decay = (float)weight_decay;
  s.neg_step_size = (float)(-(lr / bc1));
  s.bc2_sqrt = (float)sqrt(bc2);
  s.grad_scale = grad_scale;
  s.amsgrad = amsgrad ? 1 : 0;
  AdamTable tab;
  memset(&tab, 0, sizeof(tab));
  int blocks = 0, k = 0;
  for (int j = 0; j < n_tensors; ++j) {
    const cgr_adam_tensor_t& a = tensors[j];
    if (a.numel <= 0) continue;
    CGR_CHECK_ARG(a.param && a.exp_avg && a.exp_avg_sq && (!amsgrad || a.max_exp_avg_sq),
                  "cgr_peer_allreduce_adam: null pointer in tensor %d", j);
    tab.param[k] = a.param; tab.grad[k] = a.grad;        // grad = offset (in floats) of this tensor inside the arenas
    tab.m[k] = a.exp_avg; tab.v[k] = a.exp_avg_sq; tab.vmax[k] = a.max_exp_avg_sq; tab.numel[k] = a.numel;
    tab.first_block[k] = blocks;
    blocks += (int)cgr_ceil_div(a.numel, (int64_t)ADAM_CHUNK);
    ++k;
  }
  CGR_CHECK_ARG(k > 0, "cgr_peer_allreduce_adam: nothing to update");
  tab.first_block[k] = blocks;
  tab.n = k;
  PeerTable pt;
  memset(&pt, 0, sizeof(pt));
  for (int r = 0; r < world; ++r) {
    CGR_CHECK_ARG(peer_arenas[r] && peer_flags[r], "cgr_peer_allreduce_adam: null peer pointer");
    pt.arena[r] = peer_arenas[r]; pt.flags[r] = peer_flags[r];
  }
  pt.world = world; pt.rank = rank; pt.step = sync_step;
  static const double timeout_s = getenv("CGR_PEER_TIMEOUT_S") ? atof(getenv("CGR_PEER_TIMEOUT_S")) : 600.0;
  pt.timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 600.0) * 1e9);
  cgr_note_launch("peer_allreduce_adam", st, 1);
  if (peer_reduced) {                                  // two-shot
    CGR_CHECK_ARG(arena_floats > 0 && (arena_floats & 3) == 0, "cgr_peer_allreduce_adam: arena length must be a multiple of 4");
    for (int r = 0; r < world; ++r) {
      CGR_CHECK_ARG(peer_reduced[r], "cgr_peer_allreduce_adam: null reduced-slice pointer");
      pt.reduced[r] = peer_reduced[r];
    }
    pt.total = arena_floats;
    pt.slice = (cgr_ceil_div(arena_floats, (int64_t)world) + 3) / 4 * 4;
    pt.total_blocks = blocks;
    int dev = 0, sms = 0;
    CGR_CUDA(cudaGetDevice(&dev));
    CGR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // persistent: the blocks wait for each other, so all of them must be co-resident -- the grid is sized from the
    // occupancy query and launched cooperatively (the driver then schedules it only when the whole grid fits, also
    // when another stream holds SMs)
    int occ = 0;
    CGR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, peer_adam2_kernel, ADAM_THREADS, 0));
    CGR_CHECK_ARG(occ >= 1, "cgr_peer_allreduce_adam: the two-shot kernel does not fit an SM");
    if (occ > 4) occ = 4;
    const int grid = blocks < occ * sms ? blocks : occ * sms;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(ADAM_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CGR_CUDA(cudaLaunchKernelEx(&cfg, peer_adam2_kernel, tab, s, pt));
  } else {
    peer_adam_kernel<<<blocks, ADAM_THREADS, 0, st>>>(tab, s, pt);
  }
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}

extern "C" int cgr_adam_step(const cgr_adam_tensor_t* tensors, int32_t n_tensors, double lr, double beta1, double beta2,
                             double eps, double weight_decay, int64_t step, int32_t amsgrad, float grad_scale,
                             void* stream) {
  CGR_CHECK_ARG(tensors && n_tensors > 0, "cgr_adam_step: no tensors");
  CGR_CHECK_ARG(step >= 1, "cgr_adam_step: step counts from 1");
  CGR_CHECK_ARG(lr >= 0 && eps >= 0 && beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && weight_decay >= 0,
                "cgr_adam_step: hyper-parameter out of range");
  cudaStream_t st = (cudaStream_t)stream;
  // bias corrections in double, as the Python reference computes them
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars s;
  s.beta1 = (float)beta1; s.beta2 = (float)beta2;
  s.one_minus_beta1 = (float)(1.0 - beta1); s.one_minus_beta2 = (float)(1.0 - beta2);
  s.eps = (float)eps; s.weight_decay = (float)weight_decay;
  s.neg_step_size = (float)(-(lr / bc1));
  s.bc2_sqrt = (float)sqrt(bc2);
  s.grad_scale = grad_scale;
  s.amsgrad = amsgrad ? 1 : 0;
  for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_TENSORS) {      // one launch per 48 tensors (the GNN has <= 32)
    AdamTable tab;
    memset(&tab, 0, sizeof(tab));
    const int cnt = n_tensors - t0 < ADAM_MAX_TENSORS ? n_tensors - t0 : ADAM_MAX_TENSORS;
    int blocks = 0, k = 0;
    for (int j = 0; j < cnt; ++j) {
      const cgr_adam_tensor_t& a = tensors[t0 + j];
      CGR_CHECK_ARG(a.numel >= 0, "cgr_adam_step: negative size");
      if (a.numel == 0) continue;
      CGR_CHECK_ARG(a.param && a.grad && a.exp_avg && a.exp_avg_sq && (!amsgrad || a.max_exp_avg_sq),
                    "cgr_adam_step: null pointer in tensor %d", t0 + j);
      tab.param[k] = a.param; tab.grad[k] = a.grad; tab.m[k] = a.exp_avg; tab.v[k] = a.exp_avg_sq;
      tab.vmax[k] = a.max_exp_avg_sq; tab.numel[k] = a.numel;
      tab.first_block[k] = blocks;
      blocks += (int)cgr_ceil_div(a.numel, (int64_t)ADAM_CHUNK);
      ++k;
    }
    if (k == 0) continue;
    tab.first_block[k] = blocks;
    tab.n = k;
    cgr_note_launch("adam_step", st, 1);
    adam_kernel<<<blocks, ADAM_THREADS, 0, st>>>(tab, s);
    CGR_LAUNCH_CHECK();
  }
  return CGR_OK;
}
