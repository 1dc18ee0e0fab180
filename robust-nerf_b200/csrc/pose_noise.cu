// pose_noise.cu -- batched camera-pose noise initialisation and pose-error tracking.
//
// Replaces the per-pose host loops of noisy_src/noise.py:138-234 (add_noise_to_pose / add_noise_to_poses: per pose ~25
// tiny tensor ops and three `.item()` / `float()` synchronisations) and of noisy_src/noise.py:237-268 +
// noisy_src/train_pose_opt.py:232-271 (compute_pose_error in a Python loop, two synchronisations per pose) by one launch
// each, one thread per pose.  The random draws stay on the host generator in the reference's order (see
// robust_nerf_b200/noise.py); the kernels consume the raw standard-normal values.
#include "common.cuh"
#include <math.h>

namespace rn {

// noise.py:71-113 (random_rotation_matrix) + 138-191 (add_noise_to_pose) + 194-234 (add_noise_to_poses)
__global__ void pose_noise_kernel(const float* __restrict__ poses, int n, const float* __restrict__ g_angle,
                                  const float* __restrict__ g_axis, const float* __restrict__ g_trans, float rot_std_rad,
                                  float trans_std_abs, double trans_pct_over_100, float* __restrict__ out,
                                  float* __restrict__ info /*[n][2]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float P[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) P[k] = poses[(size_t)i * 16 + k];
  float rot_deg = 0.f, t_norm = 0.f;
  if (g_angle != nullptr) {
    const float a = __fmul_rn(g_angle[i], rot_std_rad);
    float ax = g_axis[i * 3], ay = g_axis[i * 3 + 1], az = g_axis[i * 3 + 2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
    ax = __fdiv_rn(ax, nrm); ay = __fdiv_rn(ay, nrm); az = __fdiv_rn(az, nrm);
    // K = [[0,-z,y],[z,0,-x],[-y,x,0]];  R = I + sin(a) K + (1 - cos(a)) K K
    const float K[9] = {0.f, -az, ay, az, 0.f, -ax, -ay, ax, 0.f};
    float KK[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        KK[r * 3 + c] = __fadd_rn(__fadd_rn(__fmul_rn(K[r * 3], K[c]), __fmul_rn(K[r * 3 + 1], K[3 + c])), __fmul_rn(K[r * 3 + 2], K[6 + c]));
    const float s = sinf(a), omc = __fsub_rn(1.f, cosf(a));
    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
      R[k] = __fadd_rn(__fadd_rn((k % 4 == 0) ? 1.f : 0.f, __fmul_rn(s, K[k])), __fmul_rn(omc, KK[k]));
    // noisy rotation = R_noise @ R_original
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        out[(size_t)i * 16 + r * 4 + c] =
            __fadd_rn(__fadd_rn(__fmul_rn(R[r * 3], P[c]), __fmul_rn(R[r * 3 + 1], P[4 + c])), __fmul_rn(R[r * 3 + 2], P[8 + c]));
    const float tr = __fadd_rn(__fadd_rn(R[0], R[4]), R[8]);
    const float cs = fminf(fmaxf(__fdiv_rn(__fsub_rn(tr, 1.f), 2.f), -1.f), 1.f);
    rot_deg = __fdiv_rn(__fmul_rn(acosf(cs), 180.f), 3.14159265358979323846f);
  } else {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) out[(size_t)i * 16 + r * 4 + c] = P[r * 4 + c];
  }
  float t[3] = {P[3], P[7], P[11]};
  if (g_trans != nullptr) {
    // noise.py:221-225: camera distance (fp32 norm -> Python float), std = distance * (pct / 100) in double, or the
    // absolute std; noise.py:116-135: randn(3) * std with std rounded to fp32
    const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(t[0], t[0]), __fmul_rn(t[1], t[1])), __fmul_rn(t[2], t[2])));
    const float std = trans_pct_over_100 > 0.0 ? (float)((double)dist * trans_pct_over_100) : trans_std_abs;
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { d[k] = __fmul_rn(g_trans[i * 3 + k], std); t[k] = __fadd_rn(t[k], d[k]); }
    t_norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  }
  out[(size_t)i * 16 + 3] = t[0]; out[(size_t)i * 16 + 7] = t[1]; out[(size_t)i * 16 + 11] = t[2];
#pragma unroll
  for (int k = 12; k < 16; ++k) out[(size_t)i * 16 + k] = P[k];
  if (info) { info[i * 2] = rot_deg; info[i * 2 + 1] = t_norm; }
}

// noise.py:237-268 (compute_pose_error), one thread per pose pair
__global__ void pose_error_kernel(const float* __restrict__ gt, const float* __restrict__ cur, int n,
                                  float* __restrict__ err /*[n][2]: rotation (deg), translation*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* G = gt + (size_t)i * 16;
  const float* C = cur + (size_t)i * 16;
  // trace(R_gt^T R_cur) = sum_c sum_k R_gt[k][c] R_cur[k][c], each diagonal entry summed over k in order
  float tr = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(G[c], C[c]), __fmul_rn(G[4 + c], C[4 + c])), __fmul_rn(G[8 + c], C[8 + c]));
    tr = (c == 0) ? d : __fadd_rn(tr, d);
  }
  const float cs = fminf(fmaxf(__fdiv_rn(__fsub_rn(tr, 1.f), 2.f), -1.f), 1.f);
  err[i * 2] = __fdiv_rn(__fmul_rn(acosf(cs), 180.f), 3.14159265358979323846f);
  const float dx = __fsub_rn(G[3], C[3]), dy = __fsub_rn(G[7], C[7]), dz = __fsub_rn(G[11], C[11]);
  err[i * 2 + 1] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

}  // namespace rn

extern "C" {

int rn_pose_noise(const float* poses, int n, const float* g_angle, const float* g_axis, const float* g_trans,
                  float rot_std_rad, float trans_std_abs, double trans_pct, float* noisy_out, float* info_out,
                  rn_stream_t stream) {
  if (n == 0) return RN_OK;
  RN_REQUIRE(poses && noisy_out && n > 0 && ((g_angle == nullptr) == (g_axis == nullptr)) && trans_pct >= 0.0);
  rn::pose_noise_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(poses, n, g_angle, g_axis, g_trans, rot_std_rad,
                                                                           trans_std_abs, trans_pct / 100.0, noisy_out, info_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_pose_errors(const float* gt_poses, const float* cur_poses, int n, float* err_out, rn_stream_t stream) {
  if (n == 0) return RN_OK;
  RN_REQUIRE(gt_poses && cur_poses && err_out && n > 0);
  rn::pose_error_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gt_poses, cur_poses, n, err_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // extern "C"
