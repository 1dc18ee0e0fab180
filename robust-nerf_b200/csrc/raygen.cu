// raygen.cu -- SE(3) exp-map pose update, ray generation, and their backward passes.
//
// Replaces (reference paths relative to the reference repo):
//   noisy_src/train_pose_opt.py:122-226  axis_angle_to_rotation_matrix / _skew_symmetric / get_poses
//   noisy_src/rays.py:17-99              get_ray_directions / get_rays
//   noisy_src/data_pose_opt.py:83-148    get_rays_from_pixels (Python loop over unique images)
//   noisy_src/data_pose_opt.py:56-76,188-198  PixelSampler bookkeeping tables
//
// All of this is tiny, latency/launch-bound work in the reference (~700 nonzero + 400 index
// launches per step); here it is one launch forward and one launch backward.  The backward is a
// deterministic per-image segmented reduction (one CTA per image, fixed reduction tree) so
// data-parallel runs are reproducible -- no float atomics.
#include "common.cuh"

namespace rn {

// Rodrigues exactly as the reference composes it: theta=|w|; small=theta<1e-6; theta<-1 if small;
// k=w/theta; K=skew(k); K2=K@K; R=I+sin*K+(1-cos)*K2; R<-I if small.
__device__ __forceinline__ void rodrigues(const float w[3], float R[9], bool& small, float& theta,
                                          float K[9], float K2[9], float& s, float& c) {
  theta = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2])));
  small = theta < 1e-6f;
  if (small) theta = 1.0f;
  const float k0 = __fdiv_rn(w[0], theta), k1 = __fdiv_rn(w[1], theta), k2 = __fdiv_rn(w[2], theta);
  K[0] = 0.f; K[1] = -k2; K[2] = k1;
  K[3] = k2;  K[4] = 0.f; K[5] = -k0;
  K[6] = -k1; K[7] = k0;  K[8] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = K[i * 3 + 0] * K[0 * 3 + j];
      a = fmaf(K[i * 3 + 1], K[1 * 3 + j], a);
      a = fmaf(K[i * 3 + 2], K[2 * 3 + j], a);
      K2[i * 3 + j] = a;
    }
  sincosf(theta, &s, &c);
  const float omc = 1.0f - c;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float I = (i % 4 == 0) ? 1.f : 0.f;
    R[i] = small ? I : __fadd_rn(__fadd_rn(I, __fmul_rn(s, K[i])), __fmul_rn(omc, K2[i]));
  }
}

// R_new = exp(w) @ R_init, t_new = t_init + dt   (train_pose_opt.py:206-218)
__device__ __forceinline__ void compose_pose(const float* __restrict__ P0, const float* w, const float* dt,
                                             bool learn_r, bool learn_t, float Rn[9], float tn[3]) {
  if (learn_r) {
    float R[9], K[9], K2[9], th, s, c; bool small;
    float wv[3] = {w[0], w[1], w[2]};
    rodrigues(wv, R, small, th, K, K2, s, c);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float a = R[i * 3 + 0] * P0[0 * 4 + j];
        a = fmaf(R[i * 3 + 1], P0[1 * 4 + j], a);
        a = fmaf(R[i * 3 + 2], P0[2 * 4 + j], a);
        Rn[i * 3 + j] = a;
      }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = P0[i * 4 + j];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) tn[i] = learn_t ? __fadd_rn(P0[i * 4 + 3], dt[i]) : P0[i * 4 + 3];
}

// dL/dw from G = dL/dR_delta (zero in the small-angle branch: reference quirk)
__device__ __forceinline__ void rodrigues_bwd(const float w[3], const float G[9], float dw[3]) {
  float R[9], K[9], K2[9], th, s, c; bool small;
  rodrigues(w, R, small, th, K, K2, s, c);
  if (small) { dw[0] = dw[1] = dw[2] = 0.f; return; }
  float d_s = 0.f, d_omc = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) { d_s = fmaf(G[i], K[i], d_s); d_omc = fmaf(G[i], K2[i], d_omc); }
  float d_th = d_s * c + d_omc * s;
  const float omc = 1.0f - c;
  // dK = s*G + (1-c)*(G K^T + K^T G)
  float dK[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = 0.f;
#pragma unroll
      for (int m = 0; m < 3; ++m) a += G[i * 3 + m] * K[j * 3 + m] + K[m * 3 + i] * G[m * 3 + j];
      dK[i * 3 + j] = s * G[i * 3 + j] + omc * a;
    }
  const float dk[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
  const float dot = dk[0] * w[0] + dk[1] * w[1] + dk[2] * w[2];
  d_th -= dot / (th * th);
#pragma unroll
  for (int i = 0; i < 3; ++i) dw[i] = dk[i] / th + d_th * w[i] / th;
}

__device__ __forceinline__ void pixel_dir(float uf, float vf, float focal, float cx, float cy, float d[3]) {
  // data_pose_opt.py:126-128: coords are cast with .long() then index the (H,W,3) direction table
  const float u = (float)(long long)uf, v = (float)(long long)vf;
  d[0] = __fdiv_rn(__fsub_rn(u, cx), focal);
  d[1] = -__fdiv_rn(__fsub_rn(v, cy), focal);
  d[2] = -1.0f;
}

// rays.py:89-97: v = sum_k dir_k * R[:,k] (mul then add, unfused), d = v/|v|
__device__ __forceinline__ float rotate_normalise(const float R[9], const float dir[3], float out[3]) {
  float v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    v[i] = __fadd_rn(__fadd_rn(__fmul_rn(dir[0], R[i * 3 + 0]), __fmul_rn(dir[1], R[i * 3 + 1])),
                     __fmul_rn(dir[2], R[i * 3 + 2]));
  const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = __fdiv_rn(v[i], n);
  return n;
}

// ------------------------------------------------------------------------------------------
__global__ void se3_poses_fwd_kernel(const float* __restrict__ P0, const float* __restrict__ rot,
                                     const float* __restrict__ trans, const int64_t* __restrict__ idx, int n,
                                     int learn_r, int learn_t, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = idx ? idx[i] : i;
  float Rn[9], tn[3];
  compose_pose(P0 + p * 16, rot + p * 3, trans + p * 3, learn_r, learn_t, Rn, tn);
  float* o = out + (int64_t)i * 16;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[r * 4 + 0] = Rn[r * 3 + 0]; o[r * 4 + 1] = Rn[r * 3 + 1]; o[r * 4 + 2] = Rn[r * 3 + 2]; o[r * 4 + 3] = tn[r];
  }
  o[12] = o[13] = o[14] = 0.f; o[15] = 1.f;
}

__device__ __forceinline__ void pose_param_grads(const float* __restrict__ P0, const float* w, const float* gP /*4x4*/,
                                                 float dw[3], float dt[3]) {
  // G = dL/dR_delta = gR_new @ R_init^T
  float G[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      G[i * 3 + j] = gP[i * 4 + 0] * P0[j * 4 + 0] + gP[i * 4 + 1] * P0[j * 4 + 1] + gP[i * 4 + 2] * P0[j * 4 + 2];
  float wv[3] = {w[0], w[1], w[2]};
  rodrigues_bwd(wv, G, dw);
  dt[0] = gP[3]; dt[1] = gP[7]; dt[2] = gP[11];
}

__global__ void se3_poses_bwd_kernel(const float* __restrict__ P0, const float* __restrict__ rot,
                                     const int64_t* __restrict__ idx, int n, const float* __restrict__ gP,
                                     float* __restrict__ d_rot, float* __restrict__ d_trans) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = idx ? idx[i] : i;
  float dw[3], dt[3];
  pose_param_grads(P0 + p * 16, rot + p * 3, gP + (int64_t)i * 16, dw, dt);
  // indices may repeat (index_select backward) -> accumulate
  if (d_rot) { atomicAdd(d_rot + p * 3 + 0, dw[0]); atomicAdd(d_rot + p * 3 + 1, dw[1]); atomicAdd(d_rot + p * 3 + 2, dw[2]); }
  if (d_trans) { atomicAdd(d_trans + p * 3 + 0, dt[0]); atomicAdd(d_trans + p * 3 + 1, dt[1]); atomicAdd(d_trans + p * 3 + 2, dt[2]); }
}

__global__ void ray_directions_kernel(int H, int W, float focal, float cx, float cy, float* __restrict__ dirs) {
  const int64_t n = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d[3];
    pixel_dir((float)(i % W), (float)(i / W), focal, cx, cy, d);
    dirs[i * 3 + 0] = d[0]; dirs[i * 3 + 1] = d[1]; dirs[i * 3 + 2] = d[2];
  }
}

__global__ void get_rays_kernel(const float* __restrict__ dirs, const float* __restrict__ c2w, int64_t n,
                                float* __restrict__ ro, float* __restrict__ rd) {
  float R[9], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { R[i * 3] = c2w[i * 4]; R[i * 3 + 1] = c2w[i * 4 + 1]; R[i * 3 + 2] = c2w[i * 4 + 2]; t[i] = c2w[i * 4 + 3]; }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d[3] = {dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]};
    float o[3];
    rotate_normalise(R, d, o);
    rd[i * 3] = o[0]; rd[i * 3 + 1] = o[1]; rd[i * 3 + 2] = o[2];
    ro[i * 3] = t[0]; ro[i * 3 + 1] = t[1]; ro[i * 3 + 2] = t[2];
  }
}

// Rays of the flat pixel range [ray0, ray0 + B) of ONE view, straight from the camera (rays.py:17-99: direction of pixel
// (u, v) = (i % W, i / W), rotated, normalised; origin = translation) plus the unit view direction render_rays derives
// from them (rendering.py:165) -- no (H, W, 3) direction table, no per-view ray tensors.  With `pose == nullptr` only the
// view directions of given rays are produced.
__global__ void view_rays_kernel(const float* __restrict__ pose, int W, float focal, float cx, float cy, int64_t ray0, int64_t B,
                                 float* __restrict__ ro, float* __restrict__ rd, float* __restrict__ vd) {
  float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, t[3] = {0.f, 0.f, 0.f};
  if (pose) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { R[i * 3] = pose[i * 4]; R[i * 3 + 1] = pose[i * 4 + 1]; R[i * 3 + 2] = pose[i * 4 + 2]; t[i] = pose[i * 4 + 3]; }
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    float d[3];
    if (pose) {
      const int64_t pix = ray0 + i;
      float dir[3];
      pixel_dir((float)(pix % W), (float)(pix / W), focal, cx, cy, dir);
      rotate_normalise(R, dir, d);
      rd[i * 3] = d[0]; rd[i * 3 + 1] = d[1]; rd[i * 3 + 2] = d[2];
      ro[i * 3] = t[0]; ro[i * 3 + 1] = t[1]; ro[i * 3 + 2] = t[2];
    } else {
      d[0] = rd[i * 3]; d[1] = rd[i * 3 + 1]; d[2] = rd[i * 3 + 2];
    }
    const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    vd[i * 3] = __fdiv_rn(d[0], n); vd[i * 3 + 1] = __fdiv_rn(d[1], n); vd[i * 3 + 2] = __fdiv_rn(d[2], n);
  }
}
int launch_view_rays(const float* pose, int W, float focal, float cx, float cy, int64_t ray0, int64_t B, float* ro, float* rd,
                     float* vd, cudaStream_t st) {
  view_rays_kernel<<<grid_for(B, 256), 256, 0, st>>>(pose, W, focal, cx, cy, ray0, B, ro, rd, vd);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

// fixed-order block reduction of NV values per thread; result valid in thread 0
template <int NV, int THREADS>
__device__ __forceinline__ void block_reduce_fixed(float (&v)[NV], float* smem /*[THREADS/32][NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float a = 0.f;
      for (int w = 0; w < THREADS / 32; ++w) a += smem[w * NV + k];
      v[k] = a;
    }
  }
}

// accumulate gradient contributions of one ray into acc[12] = dL/d(R row-major 9, t 3)
__device__ __forceinline__ void ray_pose_grad(const float R[9], const float dir[3], const float* go, const float* gd,
                                              float acc[12]) {
  float d[3];
  const float n = rotate_normalise(R, dir, d);
  const float dot = d[0] * gd[0] + d[1] * gd[1] + d[2] * gd[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float gv = (gd[i] - d[i] * dot) / n;
    acc[i * 3 + 0] += gv * dir[0]; acc[i * 3 + 1] += gv * dir[1]; acc[i * 3 + 2] += gv * dir[2];
    acc[9 + i] += go[i];
  }
}

constexpr int kBwdThreads = 256;

__global__ void get_rays_bwd_kernel(const float* __restrict__ dirs, const float* __restrict__ c2w, int64_t n,
                                    const float* __restrict__ go, const float* __restrict__ gd,
                                    float* __restrict__ g_c2w, float* __restrict__ g_dirs) {
  __shared__ float red[kBwdThreads / 32 * 12];
  float R[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) { R[i * 3] = c2w[i * 4]; R[i * 3 + 1] = c2w[i * 4 + 1]; R[i * 3 + 2] = c2w[i * 4 + 2]; }
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += kBwdThreads) {
    const float dir[3] = {dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]};
    ray_pose_grad(R, dir, go + i * 3, gd + i * 3, acc);
    if (g_dirs) {
      float d[3];
      const float nn = rotate_normalise(R, dir, d);
      const float dot = d[0] * gd[i * 3] + d[1] * gd[i * 3 + 1] + d[2] * gd[i * 3 + 2];
      float gv[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) gv[r] = (gd[i * 3 + r] - d[r] * dot) / nn;
#pragma unroll
      for (int k = 0; k < 3; ++k) g_dirs[i * 3 + k] = gv[0] * R[k] + gv[1] * R[3 + k] + gv[2] * R[6 + k];
    }
  }
  block_reduce_fixed<12, kBwdThreads>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      g_c2w[r * 4] = acc[r * 3]; g_c2w[r * 4 + 1] = acc[r * 3 + 1]; g_c2w[r * 4 + 2] = acc[r * 3 + 2]; g_c2w[r * 4 + 3] = acc[9 + r];
    }
    g_c2w[12] = g_c2w[13] = g_c2w[14] = g_c2w[15] = 0.f;
  }
}

// FUSED = true: poses come from (P0, rot, trans) through the exp map, else from `poses`.
template <bool FUSED>
__global__ void raygen_fwd_kernel(const int64_t* __restrict__ img, const float* __restrict__ uv, int64_t B,
                                  const float* __restrict__ poses, const float* __restrict__ rot,
                                  const float* __restrict__ trans, int learn_r, int learn_t, float focal, float cx,
                                  float cy, float* __restrict__ ro, float* __restrict__ rd) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = img[b];
    const float* P = poses + p * 16;
    float R[9], t[3];
    if (FUSED) {
      compose_pose(P, rot + p * 3, trans + p * 3, learn_r, learn_t, R, t);
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) { R[i * 3] = P[i * 4]; R[i * 3 + 1] = P[i * 4 + 1]; R[i * 3 + 2] = P[i * 4 + 2]; t[i] = P[i * 4 + 3]; }
    }
    float dir[3], d[3];
    pixel_dir(uv[b * 2], uv[b * 2 + 1], focal, cx, cy, dir);
    rotate_normalise(R, dir, d);
    rd[b * 3] = d[0]; rd[b * 3 + 1] = d[1]; rd[b * 3 + 2] = d[2];
    ro[b * 3] = t[0]; ro[b * 3 + 1] = t[1]; ro[b * 3 + 2] = t[2];
  }
}

// one CTA per image: scan the batch, reduce this image's rays in a fixed order
template <bool FUSED>
__global__ void __launch_bounds__(kBwdThreads)
raygen_bwd_kernel(const int64_t* __restrict__ img, const float* __restrict__ uv, int64_t B,
                  const float* __restrict__ poses, const float* __restrict__ rot, int learn_r, float focal, float cx,
                  float cy, const float* __restrict__ go, const float* __restrict__ gd, float* __restrict__ g_poses,
                  float* __restrict__ d_rot, float* __restrict__ d_trans) {
  __shared__ float red[kBwdThreads / 32 * 12];
  const int64_t p = blockIdx.x;
  const float* P = poses + p * 16;
  float R[9], t[3];
  if (FUSED) {
    const float zero[3] = {0.f, 0.f, 0.f};
    compose_pose(P, rot + p * 3, zero, learn_r, 0, R, t);
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) { R[i * 3] = P[i * 4]; R[i * 3 + 1] = P[i * 4 + 1]; R[i * 3 + 2] = P[i * 4 + 2]; }
  }
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += kBwdThreads) {
    if (img[b] != p) continue;
    float dir[3];
    pixel_dir(uv[b * 2], uv[b * 2 + 1], focal, cx, cy, dir);
    ray_pose_grad(R, dir, go + b * 3, gd + b * 3, acc);
  }
  block_reduce_fixed<12, kBwdThreads>(acc, red);
  if (threadIdx.x != 0) return;
  float gP[16];
#pragma unroll
  for (int r = 0; r < 3; ++r) { gP[r * 4] = acc[r * 3]; gP[r * 4 + 1] = acc[r * 3 + 1]; gP[r * 4 + 2] = acc[r * 3 + 2]; gP[r * 4 + 3] = acc[9 + r]; }
  gP[12] = gP[13] = gP[14] = gP[15] = 0.f;
  if (FUSED) {
    float dw[3] = {0.f, 0.f, 0.f}, dt[3];
    if (learn_r) pose_param_grads(P, rot + p * 3, gP, dw, dt);
    dt[0] = gP[3]; dt[1] = gP[7]; dt[2] = gP[11];
#pragma unroll
    for (int k = 0; k < 3; ++k) { d_rot[p * 3 + k] = dw[k]; d_trans[p * 3 + k] = dt[k]; }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) g_poses[p * 16 + k] = gP[k];
  }
}

__global__ void pixel_gather_kernel(const int64_t* __restrict__ flat, int64_t B, int H, int W,
                                    const float* __restrict__ images, int64_t* __restrict__ img_out,
                                    float* __restrict__ uv_out, float* __restrict__ rgb_out) {
  const int64_t hw = (int64_t)H * W;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = flat[b];
    const int64_t im = f / hw, rem = f - im * hw;
    const int64_t v = rem / W, u = rem - v * W;
    img_out[b] = im;
    uv_out[b * 2] = (float)u; uv_out[b * 2 + 1] = (float)v;
    if (images && rgb_out) {
      rgb_out[b * 3] = images[f * 3]; rgb_out[b * 3 + 1] = images[f * 3 + 1]; rgb_out[b * 3 + 2] = images[f * 3 + 2];
    }
  }
}

// uint8 image table: rgb = k / 255.0f in fp32, the reference's own conversion (data.py:134-136)
__global__ void pixel_gather_u8_kernel(const int64_t* __restrict__ flat, int64_t B, int H, int W,
                                       const uint8_t* __restrict__ images, int64_t* __restrict__ img_out,
                                       float* __restrict__ uv_out, float* __restrict__ rgb_out) {
  const int64_t hw = (int64_t)H * W;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = flat[b];
    const int64_t im = f / hw, rem = f - im * hw;
    const int64_t v = rem / W, u = rem - v * W;
    img_out[b] = im;
    uv_out[b * 2] = (float)u; uv_out[b * 2 + 1] = (float)v;
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb_out[b * 3 + k] = __fdiv_rn((float)images[f * 3 + k], 255.0f);
  }
}

}  // namespace rn

using namespace rn;

extern "C" {

int rn_se3_poses_fwd(const float* P0, const float* rot, const float* trans, const int64_t* idx, int n, int n_total,
                     int learn_r, int learn_t, float* out, rn_stream_t stream) {
  if (n == 0) return RN_OK;
  RN_REQUIRE(P0 && rot && trans && out && n >= 0 && n_total > 0);
  se3_poses_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P0, rot, trans, idx, n, learn_r, learn_t, out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_se3_poses_bwd(const float* P0, const float* rot, const int64_t* idx, int n, int n_total, const float* gP,
                     float* d_rot, float* d_trans, rn_stream_t stream) {
  if (n == 0) return RN_OK;
  RN_REQUIRE(P0 && rot && gP && n >= 0 && n_total > 0);
  se3_poses_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P0, rot, idx, n, gP, d_rot, d_trans);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_ray_directions(int H, int W, float focal, float cx, float cy, float* dirs, rn_stream_t stream) {
  RN_REQUIRE(dirs && H > 0 && W > 0 && focal != 0.f);
  ray_directions_kernel<<<grid_for((int64_t)H * W, 256), 256, 0, (cudaStream_t)stream>>>(H, W, focal, cx, cy, dirs);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_get_rays(const float* dirs, const float* c2w, int64_t n, float* ro, float* rd, rn_stream_t stream) {
  if (n == 0) return RN_OK;
  RN_REQUIRE(dirs && c2w && ro && rd && n >= 0);
  get_rays_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dirs, c2w, n, ro, rd);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_get_rays_bwd(const float* dirs, const float* c2w, int64_t n, const float* go, const float* gd, float* g_c2w,
                    float* g_dirs, rn_stream_t stream) {
  RN_REQUIRE(dirs && c2w && go && gd && g_c2w && n >= 0);
  get_rays_bwd_kernel<<<1, kBwdThreads, 0, (cudaStream_t)stream>>>(dirs, c2w, n, go, gd, g_c2w, g_dirs);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_raygen_fwd(const int64_t* img, const float* uv, int64_t B, const float* poses, int n_poses, int H, int W,
                  float focal, float cx, float cy, float* ro, float* rd, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(img && uv && poses && ro && rd && B >= 0 && n_poses > 0 && H > 0 && W > 0);
  raygen_fwd_kernel<false><<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(img, uv, B, poses, nullptr, nullptr, 0, 0,
                                                                                focal, cx, cy, ro, rd);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_raygen_bwd(const int64_t* img, const float* uv, int64_t B, const float* poses, int n_poses, int H, int W,
                  float focal, float cx, float cy, const float* go, const float* gd, float* g_poses,
                  rn_stream_t stream) {
  RN_REQUIRE(img && uv && poses && go && gd && g_poses && B >= 0 && n_poses > 0);
  raygen_bwd_kernel<false><<<n_poses, kBwdThreads, 0, (cudaStream_t)stream>>>(img, uv, B, poses, nullptr, 0, focal, cx, cy,
                                                                               go, gd, g_poses, nullptr, nullptr);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_raygen_se3_fwd(const int64_t* img, const float* uv, int64_t B, const float* P0, const float* rot,
                      const float* trans, int n_poses, int learn_r, int learn_t, int H, int W, float focal, float cx,
                      float cy, float* ro, float* rd, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(img && uv && P0 && rot && trans && ro && rd && B >= 0 && n_poses > 0 && H > 0 && W > 0);
  raygen_fwd_kernel<true><<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(img, uv, B, P0, rot, trans, learn_r, learn_t,
                                                                               focal, cx, cy, ro, rd);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_raygen_se3_bwd(const int64_t* img, const float* uv, int64_t B, const float* P0, const float* rot, int n_poses,
                      int learn_r, int H, int W, float focal, float cx, float cy, const float* go, const float* gd,
                      float* d_rot, float* d_trans, rn_stream_t stream) {
  RN_REQUIRE(img && uv && P0 && rot && go && gd && d_rot && d_trans && B >= 0 && n_poses > 0);
  raygen_bwd_kernel<true><<<n_poses, kBwdThreads, 0, (cudaStream_t)stream>>>(img, uv, B, P0, rot, learn_r, focal, cx, cy, go,
                                                                              gd, nullptr, d_rot, d_trans);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_pixel_gather(const int64_t* flat, int64_t B, int H, int W, const float* images, int64_t* img_out, float* uv_out,
                    float* rgb_out, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(flat && img_out && uv_out && B >= 0 && H > 0 && W > 0);
  pixel_gather_kernel<<<grid_for(B, 256), 256, 0, (cudaStream_t)stream>>>(flat, B, H, W, images, img_out, uv_out, rgb_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_pixel_gather_u8(const int64_t* flat, int64_t B, int H, int W, const uint8_t* images, int64_t* img_out, float* uv_out,
                       float* rgb_out, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(flat && images && img_out && uv_out && rgb_out && B >= 0 && H > 0 && W > 0);
  pixel_gather_u8_kernel<<<grid_for(B, 256), 256, 0, (cudaStream_t)stream>>>(flat, B, H, W, images, img_out, uv_out, rgb_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // extern "C"
