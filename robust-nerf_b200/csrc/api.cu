// api.cu -- library bookkeeping (version, status strings, device query) and the optimiser tail
// of the training step: clip_grad_norm_ + Adam fused over a flat parameter buffer
// (noisy_src/train.py:115-117, noisy_src/train_pose_opt.py:398-409).
#include "common.cuh"

namespace rn {

int g_last_cuda_error = 0;
int g_pdl = 0;
unsigned long long g_launch_count = 0;

int num_sms() {
  static int cached[kMaxDevices] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsDefault;
  int& c = cached[dev & (kMaxDevices - 1)];
  if (c == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) c = n;
    else return kNumSMsDefault;
  }
  return c;
}

constexpr int kMaxGroups = 8;
struct Groups { int64_t off[kMaxGroups + 1]; float max_norm[kMaxGroups]; int n; };

constexpr int kNormChunks = 64;      // CTAs per clip group

// grid (kNormChunks, n_groups): fixed-order partial sums of squares -> part[group][chunk]
__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ g, Groups gr, float gscale, float* __restrict__ part) {
  __shared__ float red[8];
  const int grp = blockIdx.y;
  const int64_t lo = gr.off[grp], hi = gr.off[grp + 1];
  const int64_t per = (hi - lo + kNormChunks - 1) / kNormChunks;
  const int64_t a = lo + per * blockIdx.x, b = (a + per < hi) ? a + per : hi;
  float acc = 0.f;
  for (int64_t i = a + threadIdx.x; i < b; i += 256) { const float v = g[i] * gscale; acc = fmaf(v, v, acc); }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    part[grp * kNormChunks + blockIdx.x] = t;
  }
}

// clip (torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6)), grads scaled in
// place) followed by torch.optim.Adam's default update.  Every CTA re-reduces the (tiny) partial
// sums in the same fixed order, so all threads see the same norm without another launch.
__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, Groups gr, const float* __restrict__ part, float* __restrict__ norms, float lr, float b1,
                 float b2, float eps, float bc1, float bc2_sqrt, const float* __restrict__ hyper, float gscale) {
  // gscale: 1 / world size under data parallelism -- the buffer holds the all-reduced SUM of the ranks' gradients and the
  // mean is taken here (and in grad_sumsq_kernel) instead of in a separate elementwise launch
  // hyper (device, optional): [lr, 1 - beta1^t, sqrt(1 - beta2^t)] -- lets a captured CUDA graph be replayed
  // with a changing learning rate / step count without re-capturing
  if (hyper) { lr = hyper[0]; bc1 = hyper[1]; bc2_sqrt = hyper[2]; }
  __shared__ float s_coef[kMaxGroups];
  if (threadIdx.x < gr.n) {
    float t = 0.f;
    for (int c = 0; c < kNormChunks; ++c) t += part[threadIdx.x * kNormChunks + c];
    const float nrm = sqrtf(t);
    if (blockIdx.x == 0) norms[threadIdx.x] = nrm;
    const float mx = gr.max_norm[threadIdx.x];
    s_coef[threadIdx.x] = mx > 0.f ? fminf(1.0f, mx / (nrm + 1e-6f)) : 1.0f;
  }
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int grp = 0;
#pragma unroll
    for (int k = 1; k < kMaxGroups; ++k) if (k < gr.n && i >= gr.off[k]) grp = k;
    const float gi = (g[i] * gscale) * s_coef[grp];
    g[i] = gi;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

}  // namespace rn

using namespace rn;

extern "C" {

int rn_version(void) { return 100; }

const char* rn_status_string(int s) {
  switch (s) {
    case RN_OK: return "ok";
    case RN_ERR_INVALID_ARG: return "invalid argument (null pointer, bad size or unsupported configuration)";
    case RN_ERR_CUDA: return "CUDA runtime error (see rn_last_cuda_error)";
    case RN_ERR_UNSUPPORTED_ARCH: return "device is not sm_100: tcgen05/TMEM/TMA kernels cannot run";
    case RN_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    default: return "unknown status";
  }
}

int rn_last_cuda_error(void) { return g_last_cuda_error; }

unsigned long long rn_launch_count(void) { return g_launch_count; }

int rn_device_sm_count(int* out) {
  RN_REQUIRE(out);
  int dev = 0, n = 0;
  RN_CUDA_CHECK(cudaGetDevice(&dev));
  RN_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out = n;
  return RN_OK;
}

int rn_clip_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      const int64_t* group_offsets_host, const float* group_max_norm_host, int n_groups, float lr,
                      float beta1, float beta2, float eps, int step, float* norms_out, const float* hyper_dev,
                      float grad_scale, rn_stream_t stream) {
  RN_REQUIRE(params && grads && exp_avg && exp_avg_sq && norms_out && group_offsets_host && group_max_norm_host);
  RN_REQUIRE(n > 0 && n_groups >= 1 && n_groups <= kMaxGroups && step >= 1);
  Groups gr{};
  gr.n = n_groups;
  for (int i = 0; i <= n_groups; ++i) gr.off[i] = group_offsets_host[i];
  for (int i = 0; i < n_groups; ++i) gr.max_norm[i] = group_max_norm_host[i];
  RN_REQUIRE(gr.off[0] == 0 && gr.off[n_groups] == n);
  cudaStream_t st = (cudaStream_t)stream;
  float* part = norms_out + kMaxGroups;     // norms_out must hold kMaxGroups + kMaxGroups*64 floats
  grad_sumsq_kernel<<<dim3(kNormChunks, n_groups), 256, 0, st>>>(grads, gr, grad_scale, part);
  RN_LAUNCH_CHECK();
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  clip_adam_kernel<<<grid_for(n, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, gr, part, norms_out, lr, beta1,
                                                     beta2, eps, bc1, bc2_sqrt, hyper_dev, grad_scale);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // extern "C"
