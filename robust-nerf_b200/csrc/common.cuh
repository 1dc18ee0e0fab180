// common.cuh -- shared helpers for the rnerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/rnerf_b200.h"

namespace rn {

extern int g_last_cuda_error;
extern unsigned long long g_launch_count;   // kernels launched through this library (bench.py: gpu_launches)

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return RN_ERR_CUDA;
}

#define RN_CUDA_CHECK(expr)                          \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return rn::cuda_fail(_e); \
  } while (0)

#define RN_LAUNCH_CHECK()          \
  do {                             \
    ++rn::g_launch_count;          \
    RN_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

#define RN_REQUIRE(cond)                     \
  do {                                       \
    if (!(cond)) return RN_ERR_INVALID_ARG;  \
  } while (0)

constexpr int kNumSMsDefault = 148;
int num_sms();        // of the CURRENT device (cached per device)
constexpr int kMaxDevices = 64;
// One-time per-DEVICE setup (cudaFuncSetAttribute and friends are per device): true the first time it is called with
// this mask while a given device is current.  Hosts call the library from one thread per device or from one thread
// switching devices; the mask is a plain word because the setup it guards is idempotent.
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & (kMaxDevices - 1));
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid size for a grid-stride kernel: a multiple of the SM count, capped by the work.
inline int grid_for(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t need = ceil_div(work_items, threads);
  int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  int64_t g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}

// ---- programmatic dependent launch (PDL) ----
// rn_set_flag(5, 1): the GEMM-family kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so a
// kernel's CTAs may be scheduled (barrier init, TMEM allocation, tensor-map prefetch) as soon as the previous kernel's
// CTAs leave their SMs, instead of after the whole grid has drained and a launch latency has passed.  Those kernels call
// pdl_wait() before their first read of global memory (a no-op when launched without the attribute) and
// pdl_launch_dependents() on entry.  MEASURED (profiles/r02_ab_log.md, block 15): 1-3 % SLOWER for the training step --
// the persistent kernels fill every SM until their last tile, so there is nothing to overlap and the early CTAs only add
// scheduling work; the flag stays off.
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- network geometry (reference default ModelConfig) ----
constexpr int kHidden = 256;
constexpr int kPosFreqs = 10, kDirFreqs = 4;
constexpr int kPosDim = 63, kDirDim = 27;
constexpr int kPosPad = 64;       // x_enc padded to one 64-wide K chunk
constexpr int kCatW = 320;        // [x_enc(64) | h(256)] and [feat(256) | d_enc(27) | 0]
constexpr int kHalf = 128;

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

}  // namespace rn
