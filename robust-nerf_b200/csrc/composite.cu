// composite.cu -- alpha compositing (volume rendering) forward and backward, MSE loss.
//
// Replaces noisy_src/rendering.py:20-116 (raw2outputs: sub, cat, norm, mul, relu, exp, cumprod,
// sum x3 and the autograd graph behind them) and the loss lines of noisy_src/train.py:88-99.
//
// Warp-per-ray, register-resident transmittance scan: a ray is walked in rounds of 32 samples
// (lane = sample, so every global access is a fully coalesced 128 B / 512 B request); inside a
// round the transmittance is a multiplicative warp scan, across rounds it is one carried register.
// The backward kernel keeps alpha, T, colour and depth of every round in registers between its
// forward sweep and its reverse (suffix-sum) sweep -- nothing is re-read from HBM.
// Optional early termination (t_min > 0): once the carried transmittance drops below t_min the
// remaining samples get weight 0 without evaluating exp().
//
// HBM-bound.  Algorithmic bytes per ray and pass: forward 24*S + 32,
// forward + backward 60*S + 56.
#include "common.cuh"

namespace rn {

constexpr int kCompWarps = 8;
#ifndef RN_COMP_BWD12_BLOCKS
#define RN_COMP_BWD12_BLOCKS 2      // resident CTAs asked of the 7..12-round training backward (A/B knob, scripts/ab_hbm.sh)
#endif

// Colour sigmoid of the raw (fused-head) input convention: the input is the bf16 tensor-core MLP's pre-activation (tolerance
// regime 1e-2 max-abs RGB), so ex2.approx (relative error 2^-22) and rcp.approx (1 ulp) are far inside it (the IEEE-rounded
// reciprocal __frcp_rn was measured SLOWER than the accurate version: profiles/r02_ab_log.md, block 16); three of the four
// transcendentals per sample, and the kernel is issue-bound (85 % of issue slots at S = 384).  The transmittance exponential
// keeps the accurate expf: alpha = 1 - exp(-x) cancels for small x and the fp32 contract on weights is 1e-5 relative.
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, __fadd_rn(1.0f, __expf(-x))); }

// inclusive multiplicative scan over the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
// inclusive additive suffix scan (lane l gets sum over lanes >= l)
__device__ __forceinline__ float warp_suffix_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}

struct SampleIn { float c0, c1, c2, sig, z, dist; };
struct Fetched { float a, b, c, sig, z, zn, nz; };     // raw global values of one sample, no dependent math

// issue the global loads of one sample (s < S guaranteed) in either input convention
template <bool RAW>
__device__ __forceinline__ Fetched fetch_sample(const float* __restrict__ rgb, const float* __restrict__ sigma,
                                                const float4* __restrict__ raw4, const float* __restrict__ z,
                                                const float* __restrict__ noise, int64_t base, int s, int S) {
  Fetched f;
  const int64_t i = base + s;
  if (RAW) {
    const float4 r = __ldcs(raw4 + i);
    f.a = r.x; f.b = r.y; f.c = r.z; f.sig = r.w;
  } else {
    f.a = __ldcs(rgb + i * 3); f.b = __ldcs(rgb + i * 3 + 1); f.c = __ldcs(rgb + i * 3 + 2); f.sig = __ldcs(sigma + i);
  }
  f.nz = noise ? __ldcs(noise + i) : 0.f;
  f.z = __ldg(z + i);
  f.zn = (s == S - 1) ? 0.f : __ldg(z + i + 1);
  return f;
}

template <bool RAW>
__device__ __forceinline__ SampleIn activate_sample(const Fetched& f, bool has_noise, int s, int S, float nrm) {
  SampleIn o;
  if (RAW) {
    o.c0 = sigmoidf_fast(f.a); o.c1 = sigmoidf_fast(f.b); o.c2 = sigmoidf_fast(f.c);               // model.py:181,194
  } else {
    o.c0 = f.a; o.c1 = f.b; o.c2 = f.c;
  }
  o.sig = has_noise ? __fadd_rn(f.sig, f.nz) : f.sig;                                              // rendering.py:78-80
  o.z = f.z;
  const float dz = (s == S - 1) ? 1e10f : __fsub_rn(f.zn, f.z);                                    // rendering.py:67-72
  o.dist = __fmul_rn(dz, nrm);                                                                     // rendering.py:75
  return o;
}

__device__ __forceinline__ float ray_norm(const float* __restrict__ rd, int64_t b) {
  const float x = rd[b * 3], y = rd[b * 3 + 1], zc = rd[b * 3 + 2];
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(zc, zc)));
}

template <bool RAW>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ sigma, const float4* __restrict__ raw4,
                     const float* __restrict__ z, const float* __restrict__ rd, const float* __restrict__ noise,
                     int64_t B, int S, int white, float t_min, float* __restrict__ rgb_map,
                     float* __restrict__ depth_map, float* __restrict__ acc_map, float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = blockIdx.x * (int64_t)kCompWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  const int rounds = (S + 31) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    const float nrm = ray_norm(rd, b);
    const int64_t base = b * S;
    float carry = 1.0f, a0 = 0.f, a1 = 0.f, a2 = 0.f, ad = 0.f, aw = 0.f;
    int r = 0;
    Fetched cur{};
    if (lane < S) cur = fetch_sample<RAW>(rgb, sigma, raw4, z, noise, base, lane, S);
    for (; r < rounds; ++r) {
      if (t_min > 0.f && carry < t_min) break;                       // early termination (warp-uniform)
      const int s = r * 32 + lane;
      const bool valid = s < S;
      // software pipelining: the next round's loads are in flight while this round is scanned
      Fetched nxt{};
      if (s + 32 < S) nxt = fetch_sample<RAW>(rgb, sigma, raw4, z, noise, base, s + 32, S);
      float alpha = 0.f, t = 1.f;
      SampleIn in{};
      if (valid) {
        in = activate_sample<RAW>(cur, noise != nullptr, s, S, nrm);
        alpha = __fsub_rn(1.0f, expf(-__fmul_rn(fmaxf(in.sig, 0.f), in.dist)));                    // rendering.py:83
        t = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);                                             // rendering.py:90
      }
      cur = nxt;
      const float incl = warp_scan_mul(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      const float w = alpha * T;                                                                   // rendering.py:96
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      if (valid) {
        __stcs(weights + base + s, w);
        a0 = fmaf(w, in.c0, a0); a1 = fmaf(w, in.c1, a1); a2 = fmaf(w, in.c2, a2);
        ad = fmaf(w, in.z, ad); aw += w;
      }
    }
    for (; r < rounds; ++r) {                                        // terminated early: zero the tail
      const int s = r * 32 + lane;
      if (s < S) __stcs(weights + base + s, 0.f);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); ad = warp_sum(ad); aw = warp_sum(aw);
    if (lane == 0) {
      if (white) { const float bg = __fsub_rn(1.0f, aw); a0 = __fadd_rn(a0, bg); a1 = __fadd_rn(a1, bg); a2 = __fadd_rn(a2, bg); }
      rgb_map[b * 3] = a0; rgb_map[b * 3 + 1] = a1; rgb_map[b * 3 + 2] = a2;
      depth_map[b] = ad; acc_map[b] = aw;
    }
  }
}

// LEAN: the training configuration (gradient from rgb_map only: no noise, no depth / weight gradients in, no ray-direction
// gradient out).  Three of the ten per-round register arrays disappear (noise, depth, sigma -- the sigma > 0 test is
// folded into the stored interval length), which takes the S = 384 instance from 151 to <= 128 registers and from one
// to two resident CTAs per SM (profiles/r01_ncu_full_hbm_kernels.raw.csv: the kernel is latency-bound at 12 % warps active).
template <bool RAW, int MAXR, bool LEAN>
__global__ void __launch_bounds__(kCompWarps * 32, LEAN ? (MAXR <= 4 ? 4 : (MAXR <= 6 ? 3 : (MAXR <= 12 ? 2 : 1))) : 1)
composite_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ sigma, const float4* __restrict__ raw4,
                     const float* __restrict__ z, const float* __restrict__ rd, const float* __restrict__ noise,
                     int64_t B, int S, int white, const float* __restrict__ g_map, const float* __restrict__ g_depth,
                     const float* __restrict__ g_acc, const float* __restrict__ g_w, float* __restrict__ d_rgb,
                     float* __restrict__ d_sigma, float4* __restrict__ d_raw4, float* __restrict__ d_rd) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = blockIdx.x * (int64_t)kCompWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  const int rounds = (S + 31) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    const float nrm = ray_norm(rd, b);
    const int64_t base = b * S;
    const float gm0 = g_map[b * 3], gm1 = g_map[b * 3 + 1], gm2 = g_map[b * 3 + 2];
    const float gD = g_depth ? g_depth[b] : 0.f;
    const float gconst = (g_acc ? g_acc[b] : 0.f) - (white ? (gm0 + gm1 + gm2) : 0.f);
    float al[MAXR], Tn[MAXR], c0[MAXR], c1[MAXR], c2[MAXR], sg[MAXR], ds[MAXR], G[MAXR], zz[MAXR], nzv[MAXR];
    // ---- phase 0: issue every global load of the ray before any dependent math ----
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      c0[r] = c1[r] = c2[r] = 0.f; sg[r] = 0.f; ds[r] = 0.f; G[r] = 0.f; zz[r] = 0.f; nzv[r] = 0.f;
      const int s = r * 32 + lane;
      if (r < rounds && s < S) {
        const Fetched f = fetch_sample<RAW>(rgb, sigma, raw4, z, LEAN ? nullptr : noise, base, s, S);
        c0[r] = f.a; c1[r] = f.b; c2[r] = f.c; sg[r] = f.sig;
        if (LEAN) {
          ds[r] = (s == S - 1) ? 1e10f : __fsub_rn(f.zn, f.z);      // interval length dz (rendering.py:67-72)
        } else {
          zz[r] = f.z; ds[r] = f.zn; nzv[r] = f.nz;
          G[r] = g_w ? __ldcs(g_w + base + s) : 0.f;
        }
      }
    }
    float carry = 1.0f;
    // ---- forward sweep: alpha, incoming transmittance, dL/dw per sample ----
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      al[r] = 0.f; Tn[r] = 0.f;
      if (r < rounds) {
        const int s = r * 32 + lane;
        const bool valid = s < S;
        float t = 1.f;
        if (valid && LEAN) {
          if (RAW) { c0[r] = sigmoidf_fast(c0[r]); c1[r] = sigmoidf_fast(c1[r]); c2[r] = sigmoidf_fast(c2[r]); }
          const float dist = __fmul_rn(ds[r], nrm);                                                // rendering.py:75
          al[r] = __fsub_rn(1.0f, expf(-__fmul_rn(fmaxf(sg[r], 0.f), dist)));
          t = __fadd_rn(__fsub_rn(1.0f, al[r]), 1e-10f);
          ds[r] = (sg[r] > 0.f) ? dist : 0.f;             // d relu(sigma) folded in: sigma itself is not needed again
          G[r] = gm0 * c0[r] + gm1 * c1[r] + gm2 * c2[r] + gconst;
        } else if (valid) {
          Fetched f;
          f.a = c0[r]; f.b = c1[r]; f.c = c2[r]; f.sig = sg[r]; f.z = zz[r]; f.zn = ds[r]; f.nz = nzv[r];
          const SampleIn in = activate_sample<RAW>(f, noise != nullptr, s, S, nrm);
          al[r] = __fsub_rn(1.0f, expf(-__fmul_rn(fmaxf(in.sig, 0.f), in.dist)));
          t = __fadd_rn(__fsub_rn(1.0f, al[r]), 1e-10f);
          c0[r] = in.c0; c1[r] = in.c1; c2[r] = in.c2; sg[r] = in.sig; ds[r] = in.dist;
          G[r] = gm0 * in.c0 + gm1 * in.c1 + gm2 * in.c2 + gD * in.z + gconst + G[r];
        }
        const float incl = warp_scan_mul(t, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        Tn[r] = carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    // ---- reverse sweep: R_s = sum_{k>s} G_k w_k ; dL/dalpha = G T - R / t ----
    float suffix = 0.f, dn = 0.f;
#pragma unroll
    for (int r = MAXR - 1; r >= 0; --r) {
      if (r < rounds) {
        const int s = r * 32 + lane;
        const bool valid = s < S;
        const float w = al[r] * Tn[r];
        const float gw = valid ? G[r] * w : 0.f;
        const float incl = warp_suffix_sum(gw, lane);
        const float R = suffix + (incl - gw);
        suffix += __shfl_sync(0xffffffffu, incl, 0);
        if (valid) {
          const float t = __fadd_rn(__fsub_rn(1.0f, al[r]), 1e-10f);
          const float d_alpha = G[r] * Tn[r] - R / t;
          const float one_m = 1.0f - al[r];
          const float dsig = LEAN ? d_alpha * ds[r] * one_m : ((sg[r] > 0.f) ? d_alpha * ds[r] * one_m : 0.f);
          // dL/ddist * dz  (dz = dist / |d|)
          if (!LEAN) dn += d_alpha * fmaxf(sg[r], 0.f) * one_m * (ds[r] / nrm);
          const float g0 = w * gm0, g1 = w * gm1, g2 = w * gm2;
          if (RAW) {
            float4 o;
            o.x = g0 * c0[r] * (1.0f - c0[r]); o.y = g1 * c1[r] * (1.0f - c1[r]); o.z = g2 * c2[r] * (1.0f - c2[r]);
            o.w = dsig;
            __stcs(d_raw4 + base + s, o);
          } else {
            __stcs(d_rgb + (base + s) * 3, g0); __stcs(d_rgb + (base + s) * 3 + 1, g1); __stcs(d_rgb + (base + s) * 3 + 2, g2);
            __stcs(d_sigma + base + s, dsig);
          }
        }
      }
    }
    if (!LEAN && d_rd) {
      dn = warp_sum(dn);
      if (lane < 3) d_rd[b * 3 + lane] = dn * rd[b * 3 + lane] / nrm;
    }
  }
}

// The training configuration of the backward (gradient from rgb_map only: no noise, no depth / weight gradients in, no
// ray-direction gradient out) as its own kernel, built to keep more rays in flight per SM (the general kernel holds ten
// per-round register arrays; at S = 384 that is 151 registers and ONE resident CTA):
//   * per round only alpha, the three colours and the masked interval length stay in registers (5 arrays); dL/dw is
//     re-evaluated from the colours, the incoming transmittance is re-scanned from alpha in the reverse sweep with the
//     round's carry parked in lane r of one register -- same operations in the same order, so the values are those of
//     the forward sweep bit for bit;
//   * FULL = the ray fills MAXR rounds exactly (64, 128, 192, 384 samples: every benchmark configuration): no `valid`
//     predicates and no run-time trip counts, so the unrolled rounds are straight-line code and the scans / exponentials
//     of different rounds interleave (the dependent SHFL chains of one round at a time left the schedulers idle).
template <bool RAW, int MAXR, bool FULL>
__global__ void __launch_bounds__(kCompWarps * 32, MAXR <= 4 ? 4 : (MAXR <= 6 ? 3 : (MAXR <= 12 ? RN_COMP_BWD12_BLOCKS : 1)))
composite_bwd_lean_kernel(const float* __restrict__ rgb, const float* __restrict__ sigma, const float4* __restrict__ raw4,
                          const float* __restrict__ z, const float* __restrict__ rd, int64_t B, int S, int white,
                          const float* __restrict__ g_map, float* __restrict__ d_rgb, float* __restrict__ d_sigma,
                          float4* __restrict__ d_raw4) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = blockIdx.x * (int64_t)kCompWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  const int rounds = FULL ? MAXR : ((S + 31) >> 5);
  for (int64_t b = warp; b < B; b += nwarps) {
    const float nrm = ray_norm(rd, b);
    const int64_t base = b * S;
    const float gm0 = g_map[b * 3], gm1 = g_map[b * 3 + 1], gm2 = g_map[b * 3 + 2];
    const float gconst = -(white ? (gm0 + gm1 + gm2) : 0.f);
    float al[MAXR], c0[MAXR], c1[MAXR], c2[MAXR], ds[MAXR];
    // ---- phase 0: issue every global load of the ray before any dependent math (sigma parks in al) ----
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      c0[r] = c1[r] = c2[r] = 0.f; al[r] = 0.f; ds[r] = 0.f;
      const int s = r * 32 + lane;
      if (FULL || (r < rounds && s < S)) {
        const Fetched f = fetch_sample<RAW>(rgb, sigma, raw4, z, nullptr, base, s, S);
        c0[r] = f.a; c1[r] = f.b; c2[r] = f.c; al[r] = f.sig;
        ds[r] = (s == S - 1) ? 1e10f : __fsub_rn(f.zn, f.z);        // interval length dz (rendering.py:67-72)
      }
    }
    float carry = 1.0f, parked = 1.0f;
    // ---- forward sweep: alpha per sample, carried transmittance per round ----
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      if (FULL || r < rounds) {
        const int s = r * 32 + lane;
        const bool valid = FULL || s < S;
        float t = 1.f;
        if (valid) {
          if (RAW) { c0[r] = sigmoidf_fast(c0[r]); c1[r] = sigmoidf_fast(c1[r]); c2[r] = sigmoidf_fast(c2[r]); }
          const float sg = al[r];
          const float dist = __fmul_rn(ds[r], nrm);                                                // rendering.py:75
          al[r] = __fsub_rn(1.0f, expf(-__fmul_rn(fmaxf(sg, 0.f), dist)));
          t = __fadd_rn(__fsub_rn(1.0f, al[r]), 1e-10f);
          ds[r] = (sg > 0.f) ? dist : 0.f;                // d relu(sigma) folded in: sigma itself is not needed again
        }
        if (lane == r) parked = carry;                    // transmittance entering round r
        const float incl = warp_scan_mul(t, lane);
        carry *= __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    // ---- reverse sweep: R_s = sum_{k>s} G_k w_k ; dL/dalpha = G T - R / t ----
    float suffix = 0.f;
#pragma unroll
    for (int r = MAXR - 1; r >= 0; --r) {
      if (FULL || r < rounds) {
        const int s = r * 32 + lane;
        const bool valid = FULL || s < S;
        const float t = valid ? __fadd_rn(__fsub_rn(1.0f, al[r]), 1e-10f) : 1.f;
        const float incl_t = warp_scan_mul(t, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl_t, 1);
        if (lane == 0) excl = 1.0f;
        const float Tn = __shfl_sync(0xffffffffu, parked, r) * excl;
        const float G = gm0 * c0[r] + gm1 * c1[r] + gm2 * c2[r] + gconst;
        const float w = al[r] * Tn;
        const float gw = valid ? G * w : 0.f;
        const float incl = warp_suffix_sum(gw, lane);
        const float R = suffix + (incl - gw);
        suffix += __shfl_sync(0xffffffffu, incl, 0);
        if (valid) {
          const float d_alpha = G * Tn - R / t;
          const float one_m = 1.0f - al[r];
          const float dsig = d_alpha * ds[r] * one_m;
          const float g0 = w * gm0, g1 = w * gm1, g2 = w * gm2;
          if (RAW) {
            float4 o;
            o.x = g0 * c0[r] * (1.0f - c0[r]); o.y = g1 * c1[r] * (1.0f - c1[r]); o.z = g2 * c2[r] * (1.0f - c2[r]);
            o.w = dsig;
            __stcs(d_raw4 + base + s, o);
          } else {
            __stcs(d_rgb + (base + s) * 3, g0); __stcs(d_rgb + (base + s) * 3 + 1, g1); __stcs(d_rgb + (base + s) * 3 + 2, g2);
            __stcs(d_sigma + base + s, dsig);
          }
        }
      }
    }
  }
}

// train.py:89-99: loss += mean((rgb_map-target)^2); g = 2*(rgb_map-target)/(3B)*scale. One CTA,
// fixed reduction order (deterministic); B*3 elements is tiny (12,288 at B=4096).
__global__ void __launch_bounds__(1024)
mse_kernel(const float* __restrict__ rgb_map, const float* __restrict__ target, int64_t n, float scale,
           float* __restrict__ loss_out, float* __restrict__ g) {
  __shared__ float red[32];
  float acc = 0.f;
  const float inv = 1.0f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float d = rgb_map[i] - target[i];
    acc = fmaf(d, d, acc);
    if (g) g[i] = 2.0f * d * inv * scale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += red[w];
    loss_out[0] += t * inv;
  }
}

// The training loss of train.py:88-99 in ONE launch: loss = mean((rgb_coarse - t)^2) [+ mean((rgb_fine - t)^2)], written
// (not accumulated) as out[0] = total, out[1] = coarse, out[2] = fine, with both gradients 2 (rgb - t) / (3B).  One CTA,
// fixed reduction order (deterministic).  Replaces two mse_kernel launches, two zero fills, an add and two multiplies.
__global__ void __launch_bounds__(1024)
mse2_kernel(const float* __restrict__ rgb_c, const float* __restrict__ rgb_f, const float* __restrict__ target, int64_t n,
            float* __restrict__ out, float* __restrict__ g_c, float* __restrict__ g_f) {
  __shared__ float red[2][32];
  float ac = 0.f, af = 0.f;
  const float inv = 1.0f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float t = target[i];
    const float dc = rgb_c[i] - t;
    ac = fmaf(dc, dc, ac);
    g_c[i] = 2.0f * dc * inv;
    if (rgb_f) {
      const float df = rgb_f[i] - t;
      af = fmaf(df, df, af);
      g_f[i] = 2.0f * df * inv;
    }
  }
  ac = warp_sum(ac); af = warp_sum(af);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ac; red[1][threadIdx.x >> 5] = af; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tc = 0.f, tf = 0.f;
    for (int w = 0; w < 32; ++w) { tc += red[0][w]; tf += red[1][w]; }
    tc *= inv; tf *= inv;
    out[0] = tc + tf; out[1] = tc; out[2] = tf;
  }
}

template <bool RAW>
static int launch_bwd(int rounds, dim3 grid, cudaStream_t st, const float* rgb, const float* sigma, const float4* raw4,
                      const float* z, const float* rd, const float* noise, int64_t B, int S, int white, const float* g_map,
                      const float* g_depth, const float* g_acc, const float* g_w, float* d_rgb, float* d_sigma,
                      float4* d_raw4, float* d_rd) {
  const bool lean = !noise && !g_depth && !g_w && !d_rd;
  // persistent grid-stride kernel: one wave of exactly the CTAs that are resident at this instantiation's register count
  auto launch = [&](auto kern) {
    int per_sm = 1;                        // host-side query (no device work; legal during stream capture)
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCompWarps * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t g = (int64_t)num_sms() * per_sm;
    if (g > grid.x || grid.x <= 2 * g) g = grid.x;      // small batches: one CTA per 8 rays, let the hardware schedule the waves
    kern<<<(unsigned)g, kCompWarps * 32, 0, st>>>(rgb, sigma, raw4, z, rd, noise, B, S, white, g_map, g_depth, g_acc, g_w, d_rgb,
                                                  d_sigma, d_raw4, d_rd);
  };
  auto launch_lean = [&](auto kern) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCompWarps * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t g = (int64_t)num_sms() * per_sm;
    if (g > grid.x || grid.x <= 2 * g) g = grid.x;
    kern<<<(unsigned)g, kCompWarps * 32, 0, st>>>(rgb, sigma, raw4, z, rd, B, S, white, g_map, d_rgb, d_sigma, d_raw4);
  };
#define RN_BWD_CASE(R)                                                     \
  do {                                                                     \
    if (lean && S == (R) * 32) launch_lean(composite_bwd_lean_kernel<RAW, R, true>);   \
    else if (lean) launch_lean(composite_bwd_lean_kernel<RAW, R, false>);  \
    else launch(composite_bwd_kernel<RAW, R, false>);                      \
  } while (0)
  if (rounds <= 2) RN_BWD_CASE(2);
  else if (rounds <= 4) RN_BWD_CASE(4);
  else if (rounds <= 6) RN_BWD_CASE(6);
  else if (rounds <= 8) RN_BWD_CASE(8);
  else if (rounds <= 12) RN_BWD_CASE(12);
  else if (rounds <= 16) RN_BWD_CASE(16);
  else return RN_ERR_INVALID_ARG;
#undef RN_BWD_CASE
  return RN_OK;
}

}  // namespace rn

using namespace rn;

extern "C" {

int rn_composite_fwd(const float* rgb, const float* sigma, const float* raw4, const float* z, const float* rd,
                     const float* noise, int64_t B, int S, int white, float t_min, float* rgb_map, float* depth_map,
                     float* acc_map, float* weights, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(z && rd && rgb_map && depth_map && acc_map && weights && B >= 0 && S >= 1);
  RN_REQUIRE((raw4 != nullptr) != (rgb != nullptr && sigma != nullptr));
  const int grid = (int)(ceil_div(B, kCompWarps) < (int64_t)num_sms() * 8 ? ceil_div(B, kCompWarps) : (int64_t)num_sms() * 8);
  if (raw4)
    composite_fwd_kernel<true><<<grid, kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        nullptr, nullptr, (const float4*)raw4, z, rd, noise, B, S, white, t_min, rgb_map, depth_map, acc_map, weights);
  else
    composite_fwd_kernel<false><<<grid, kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        rgb, sigma, nullptr, z, rd, noise, B, S, white, t_min, rgb_map, depth_map, acc_map, weights);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_composite_bwd(const float* rgb, const float* sigma, const float* raw4, const float* z, const float* rd,
                     const float* noise, int64_t B, int S, int white, const float* g_map, const float* g_depth,
                     const float* g_acc, const float* g_w, float* d_rgb, float* d_sigma, float* d_raw4, float* d_rd,
                     rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(z && rd && g_map && B >= 0 && S >= 1);
  RN_REQUIRE((raw4 != nullptr) != (rgb != nullptr && sigma != nullptr));
  RN_REQUIRE(raw4 ? (d_raw4 != nullptr) : (d_rgb != nullptr && d_sigma != nullptr));
  const int rounds = (S + 31) / 32;
  const int grid = (int)(ceil_div(B, kCompWarps) < (int64_t)num_sms() * 8 ? ceil_div(B, kCompWarps) : (int64_t)num_sms() * 8);
  int rc;
  if (raw4)
    rc = launch_bwd<true>(rounds, dim3(grid), (cudaStream_t)stream, nullptr, nullptr, (const float4*)raw4, z, rd, noise, B, S,
                          white, g_map, g_depth, g_acc, g_w, nullptr, nullptr, (float4*)d_raw4, d_rd);
  else
    rc = launch_bwd<false>(rounds, dim3(grid), (cudaStream_t)stream, rgb, sigma, nullptr, z, rd, noise, B, S, white, g_map,
                           g_depth, g_acc, g_w, d_rgb, d_sigma, nullptr, d_rd);
  if (rc != RN_OK) return rc;
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_mse2_loss_fwd_bwd(const float* rgb_coarse, const float* rgb_fine, const float* target, int64_t B, float* loss_out,
                         float* g_coarse, float* g_fine, rn_stream_t stream) {
  RN_REQUIRE(rgb_coarse && target && loss_out && g_coarse && B > 0 && ((rgb_fine == nullptr) == (g_fine == nullptr)));
  mse2_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rgb_coarse, rgb_fine, target, B * 3, loss_out, g_coarse, g_fine);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_mse_loss_fwd_bwd(const float* rgb_map, const float* target, int64_t B, float loss_scale, float* loss_out,
                        float* g_rgb_map, rn_stream_t stream) {
  RN_REQUIRE(rgb_map && target && loss_out && B > 0);
  mse_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rgb_map, target, B * 3, loss_scale, loss_out, g_rgb_map);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // extern "C"
