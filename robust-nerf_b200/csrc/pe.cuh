// pe.cuh -- positional-encoding features for the tensor-core operand (noisy_src/model.py:58-80: x, then per frequency
// sin(2^k xyz), cos(2^k xyz); no pi factor), shared by the in-kernel encoder of the forward chain (chain_pair.cu) and
// the per-ray view-direction projection (encode.cu).
//
// The features feed a bf16 operand (8 mantissa bits), so they are produced with FOUR accurate sincosf per coordinate
// (k = 0, 3, 6, 9: Cody-Waite range reduction, arguments reach 3,000 rad) and two angle doublings after each
//     sin 2a = 2 sin a cos a,   cos 2a = 1 - 2 sin^2 a
// instead of ten accurate calls.  Absolute error of the doubled values: <= 5e-7 (measured against float64 on 200,000
// points in [-6, 6]: profiles/r02_pe_recurrence.md), i.e. 1e-4 of the bf16 rounding step; one feature in 10^4 lands on
// the neighbouring bf16 value.  The fp32 PositionalEncoding.forward of the API (posenc_fwd_kernel) keeps ten accurate calls.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rn {

// feat[0 .. 3 + 6L) for one 3-vector
template <int L>
__device__ __forceinline__ void pe_features_fast(const float x[3], float* feat) {
  feat[0] = x[0]; feat[1] = x[1]; feat[2] = x[2];
  float s[3], c[3];
#pragma unroll
  for (int k = 0; k < L; ++k) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (k % 3 == 0) {
        sincosf(__fmul_rn((float)(1 << k), x[a]), &s[a], &c[a]);
      } else {
        const float s2 = __fadd_rn(s[a], s[a]);
        const float cn = __fsub_rn(1.0f, __fmul_rn(s2, s[a]));
        s[a] = __fmul_rn(s2, c[a]);
        c[a] = cn;
      }
      feat[3 + 6 * k + a] = s[a];
      feat[3 + 6 * k + 3 + a] = c[a];
    }
  }
}

__device__ __forceinline__ uint32_t pe_pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));      // low half = a, high half = b
  return r;
}

}  // namespace rn
