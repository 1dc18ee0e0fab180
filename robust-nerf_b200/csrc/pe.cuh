// pe.cuh -- positional-encoding features for the tensor-core operand (noisy_src/model.py:58-80: x, then per frequency
// sin(2^k xyz), cos(2^k xyz); no pi factor), shared by the in-kernel encoder of the forward chain (chain_pair.cu) and
// the per-ray view-direction projection (encode.cu).
//
// The features feed a bf16 operand (8 mantissa bits), so instead of ten accurate sincosf per coordinate (Cody-Waite range
// reduction of arguments that reach 3,000 rad: ~2,000 instructions per point) the argument is reduced ONCE, in turns:
//     t = x / (2 pi) as a two-float value (t_hi + t_lo, FMA residual), so that 2^k t is exact scaling;
//     r_k = frac(2^k t_hi) + 2^k t_lo  in [-0.5, 0.5];   sin/cos(2^k x) = sin/cos(2 pi r_k)  through MUFU (sin.approx /
//     cos.approx, absolute error 2^-21.4 on [-pi, pi]).
// Measured against float64 (profiles/r02_pe_recurrence.md): absolute error <= 8e-7 at EVERY frequency, i.e. 2e-4 of the
// bf16 rounding step; 3.5 features in 10^4 land on the neighbouring bf16 value.  ~10 instructions per (frequency,
// coordinate).  The fp32 PositionalEncoding.forward of the API (posenc_fwd_kernel) keeps the accurate sincosf.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rn {

// t = x / (2 pi) as a two-float value
__device__ __forceinline__ void pe_turns(float x, float& t_hi, float& t_lo) {
  constexpr float kInv2PiHi = 0.15915494f;                 // fl32(1 / (2 pi))
  constexpr float kInv2PiLo = 6.4206382e-09f;              // fl32(1 / (2 pi) - kInv2PiHi)
  t_hi = __fmul_rn(x, kInv2PiHi);
  t_lo = __fmaf_rn(x, kInv2PiLo, __fmaf_rn(x, kInv2PiHi, -t_hi));
}
// sin / cos of f * x for f = 2^k, from the turns of x
__device__ __forceinline__ void pe_sincos_turns(float t_hi, float t_lo, float f, float& sn, float& cs) {
  constexpr float k2Pi = 6.2831855f;
  const float y = __fmul_rn(t_hi, f);                      // exact
  const float r = __fmaf_rn(t_lo, f, __fsub_rn(y, rintf(y)));
  const float ang = __fmul_rn(r, k2Pi);
  sn = __sinf(ang);
  cs = __cosf(ang);
}

// feat[0 .. 3 + 6L) for one 3-vector
template <int L>
__device__ __forceinline__ void pe_features_fast(const float x[3], float* feat) {
  feat[0] = x[0]; feat[1] = x[1]; feat[2] = x[2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float t_hi, t_lo;
    pe_turns(x[a], t_hi, t_lo);
#pragma unroll
    for (int k = 0; k < L; ++k) pe_sincos_turns(t_hi, t_lo, (float)(1 << k), feat[3 + 6 * k + a], feat[3 + 6 * k + 3 + a]);
  }
}

__device__ __forceinline__ uint32_t pe_pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));      // low half = a, high half = b
  return r;
}

}  // namespace rn
