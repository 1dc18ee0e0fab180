// metrics.cu -- batched image metrics of the evaluation loop on the GPU.
//
// Replaces noisy_src/metrics.py:15-116 (compute_mse / compute_psnr / compute_ssim: per image a host-side loop of
// 5 conv2d + ~15 elementwise launches and a .item() each) by ONE launch over a whole batch of rendered views plus a
// deterministic reduction: per image the mean squared error (PSNR = 20 log10(max) - 10 log10(mse) is a host formula)
// and the mean of the per-channel Gaussian-window SSIM map (11 x 11 window, sigma 1.5, zero padding, C1 = 0.01^2,
// C2 = 0.03^2 by default).  The window weights are the reference's own fp32 outer product g_i * g_j.
//
// One CTA = one 32 x 32 output tile of one image, all three channels: the (32+10)^2 x 3 neighbourhoods of pred and
// target are staged in shared memory (coalesced channel-last rows), each thread evaluates 4 pixels x 3 channels with the
// direct 121-tap window for the five moments, and the CTA writes two partial sums.  No atomics: bit-reproducible.
#include "common.cuh"
#include <math.h>

namespace rn {

constexpr int kWin = 11, kHalo = 5, kTile = 32, kPatch = kTile + 2 * kHalo;     // 42
constexpr int kMetThreads = 256;
// 11x11 Gaussian window, passed by value as a kernel parameter (parameters live in the constant bank): no process-global
// __constant__ symbol, so no per-device first-use copy and nothing illegal under stream capture
struct SsimWindow { float w[kWin * kWin]; };

__global__ void __launch_bounds__(kMetThreads)
image_metrics_tile_kernel(const __grid_constant__ SsimWindow win, const float* __restrict__ pred, const float* __restrict__ target, int H, int W, float C1, float C2,
                          float* __restrict__ partial /*[N][tiles][2]*/) {
  extern __shared__ float smem[];
  float* s_p = smem;                                   // [42][42][3]
  float* s_t = smem + kPatch * kPatch * 3;
  __shared__ float s_red[2][kMetThreads / 32];
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kTile - kHalo, x0 = blockIdx.x * kTile - kHalo;
  const float* P = pred + (size_t)n * H * W * 3;
  const float* T = target + (size_t)n * H * W * 3;
  for (int i = threadIdx.x; i < kPatch * kPatch * 3; i += kMetThreads) {
    const int c = i % 3, px = (i / 3) % kPatch, py = i / (3 * kPatch);
    const int y = y0 + py, x = x0 + px;
    const bool in = (y >= 0 && y < H && x >= 0 && x < W);          // zero padding (F.conv2d padding = 5)
    const size_t g = ((size_t)y * W + x) * 3 + c;
    s_p[i] = in ? __ldg(P + g) : 0.f;
    s_t[i] = in ? __ldg(T + g) : 0.f;
  }
  __syncthreads();
  float ssim_sum = 0.f, se_sum = 0.f;
  for (int o = threadIdx.x; o < kTile * kTile; o += kMetThreads) {
    const int oy = o / kTile, ox = o % kTile;
    const int y = blockIdx.y * kTile + oy, x = blockIdx.x * kTile + ox;
    if (y >= H || x >= W) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float mp = 0.f, mt = 0.f, pp = 0.f, tt = 0.f, pt = 0.f;
      for (int i = 0; i < kWin; ++i) {
        const float* rp = s_p + ((oy + i) * kPatch + ox) * 3 + c;
        const float* rt = s_t + ((oy + i) * kPatch + ox) * 3 + c;
#pragma unroll
        for (int j = 0; j < kWin; ++j) {
          const float w = win.w[i * kWin + j];
          const float a = rp[j * 3], b = rt[j * 3];
          mp = fmaf(w, a, mp); mt = fmaf(w, b, mt);
          pp = fmaf(w, a * a, pp); tt = fmaf(w, b * b, tt); pt = fmaf(w, a * b, pt);
        }
      }
      const float mpp = mp * mp, mtt = mt * mt, mpt = mp * mt;                       // metrics.py:102-104
      const float spp = pp - mpp, stt = tt - mtt, spt = pt - mpt;                     // metrics.py:107-109
      ssim_sum += ((2.f * mpt + C1) * (2.f * spt + C2)) / ((mpp + mtt + C1) * (spp + stt + C2));   // metrics.py:112-113
      const float d = s_p[((oy + kHalo) * kPatch + ox + kHalo) * 3 + c] - s_t[((oy + kHalo) * kPatch + ox + kHalo) * 3 + c];
      se_sum = fmaf(d, d, se_sum);
    }
  }
  ssim_sum = warp_sum(ssim_sum); se_sum = warp_sum(se_sum);
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = ssim_sum; s_red[1][threadIdx.x >> 5] = se_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < kMetThreads / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }     // fixed order
    const int tiles = gridDim.x * gridDim.y;
    float* o = partial + ((size_t)n * tiles + blockIdx.y * gridDim.x + blockIdx.x) * 2;
    o[0] = a; o[1] = b;
  }
}

// one warp per image: fixed-order sum of the tile partials
__global__ void image_metrics_reduce_kernel(const float* __restrict__ partial, int N, int tiles, float inv_count,
                                            float* __restrict__ mse_out, float* __restrict__ ssim_out) {
  const int n = blockIdx.x, lane = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int t = lane; t < tiles; t += 32) { a += partial[((size_t)n * tiles + t) * 2]; b += partial[((size_t)n * tiles + t) * 2 + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (lane == 0) { ssim_out[n] = (float)(a * inv_count); mse_out[n] = (float)(b * inv_count); }
}

}  // namespace rn

using namespace rn;

extern "C" {

size_t rn_image_metrics_scratch_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  const size_t tiles = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
  return (size_t)N * tiles * 2 * sizeof(float);
}

int rn_image_metrics(const float* pred, const float* target, int N, int H, int W, float C1, float C2, float* scratch,
                     float* mse_out, float* ssim_out, rn_stream_t stream) {
  if (N == 0) return RN_OK;
  RN_REQUIRE(pred && target && scratch && mse_out && ssim_out && N > 0 && H > 0 && W > 0);
  // metrics.py:85-89: g = exp(-x^2 / (2 * 1.5^2)), g /= sum(g), window = outer(g, g), all in fp32
  SsimWindow win;
  {
    float g[kWin], s = 0.f;
    for (int i = 0; i < kWin; ++i) { const float x = (float)(i - kWin / 2); g[i] = expf(-(x * x) / (2.f * 1.5f * 1.5f)); s += g[i]; }
    for (int i = 0; i < kWin; ++i) g[i] /= s;
    for (int i = 0; i < kWin; ++i) for (int j = 0; j < kWin; ++j) win.w[i * kWin + j] = g[i] * g[j];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N);
  const size_t smem = (size_t)2 * kPatch * kPatch * 3 * sizeof(float);
  image_metrics_tile_kernel<<<grid, kMetThreads, smem, st>>>(win, pred, target, H, W, C1, C2, scratch);
  RN_LAUNCH_CHECK();
  image_metrics_reduce_kernel<<<N, 32, 0, st>>>(scratch, N, (int)(grid.x * grid.y), 1.0f / ((float)H * (float)W * 3.0f), mse_out, ssim_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // extern "C"
