// encode.cu -- everything of the NeRF MLP that is not a 256-wide GEMM:
//   positional encoding (noisy_src/model.py:58-80; no pi factor, sin/cos with full range reduction),
//   fused into the producer of the first layer's bf16 operand; the sigma (256->1) and rgb (128->3)
//   heads of model.py:181,194 on CUDA cores in fp32; weight packing fp32 -> padded bf16;
//   and the backward passes of all of them.
#include "common.cuh"
#include "mlp_layout.h"
#include "pe.cuh"

namespace rn {

using namespace layout;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// 256-bit global stores (sm_100: STG.E.ENL2.256): sixteen bf16 from sixteen floats, or zeros
__device__ __forceinline__ void stg256_bf16x16(__nv_bfloat16* p, const float* f) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(pack_bf16(f[0], f[1])), "r"(pack_bf16(f[2], f[3])), "r"(pack_bf16(f[4], f[5])),
                 "r"(pack_bf16(f[6], f[7])), "r"(pack_bf16(f[8], f[9])), "r"(pack_bf16(f[10], f[11])),
                 "r"(pack_bf16(f[12], f[13])), "r"(pack_bf16(f[14], f[15]))
               : "memory");
}
__device__ __forceinline__ void stg256_zero(__nv_bfloat16* p) {
  asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(0u) : "memory");
}

// feat[0..3*(1+2L)) for one 3-vector, reference order: x, then per frequency sin(xyz), cos(xyz)
template <int L>
__device__ __forceinline__ void encode3(const float x[3], float* feat) {
  feat[0] = x[0]; feat[1] = x[1]; feat[2] = x[2];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, c;
      sincosf(__fmul_rn(f, x[a]), &s, &c);     // accurate range reduction: arguments reach ~3000 rad
      feat[3 + 6 * k + a] = s;
      feat[3 + 6 * k + 3 + a] = c;
    }
  }
}

// pts[M,3] -> XC[:, 0:64] (bf16, col 63 = 0); dirs[M/group,3] -> FD[:, 256:320] (27 features + zeros)
__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ pts, const float* __restrict__ dirs, int64_t M, int group,
              __nv_bfloat16* __restrict__ XC, int ldx, __nv_bfloat16* __restrict__ FD, int ldf) {
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    float feat[64];
    const float x[3] = {__ldcs(pts + m * 3), __ldcs(pts + m * 3 + 1), __ldcs(pts + m * 3 + 2)};
    pe_features_fast<kPosFreqs>(x, feat);       // csrc/pe.cuh: four accurate sincosf per coordinate + angle doubling
    feat[63] = 0.f;
    // one full 32-byte sector per store instruction (rows are 32-byte aligned: ld is a multiple of 16 elements)
    __nv_bfloat16* dst = XC + m * ldx;
#pragma unroll
    for (int i = 0; i < 4; ++i) stg256_bf16x16(dst + 16 * i, feat + 16 * i);
    if (FD) {
      const int64_t r = m / group;
      const float d[3] = {__ldg(dirs + r * 3), __ldg(dirs + r * 3 + 1), __ldg(dirs + r * 3 + 2)};
      float df[32];
      pe_features_fast<kDirFreqs>(d, df);
#pragma unroll
      for (int i = 27; i < 32; ++i) df[i] = 0.f;
      __nv_bfloat16* dd = FD + m * ldf + 256;
      stg256_bf16x16(dd, df);
      stg256_bf16x16(dd + 16, df + 16);
      stg256_zero(dd + 32);
      stg256_zero(dd + 48);
    }
  }
}

// View-direction branch hoisted per RAY (inference): every point of a ray shares one direction, so its contribution to
// dir_linear (model.py:187-192: cat[features, d_enc] @ W_d^T + b_d) is the same 128-vector for all of them:
//   dirvec[r][j] = b_d[j] + sum_{i<27} bf16(d_enc[r][i]) * W_d[j][256 + i]
// (bf16-rounded operands, fp32 accumulation: what the tensor core did with the d_enc K-chunk).  The forward chain adds it
// as a per-row bias of the view layer instead of streaming a 64-wide K chunk per POINT: 84 -> 60 sin/cos per point, no
// d_enc tile in HBM, one K chunk less in the view layer.  One warp per ray.
__global__ void __launch_bounds__(256)
dir_bias_kernel(const float* __restrict__ dirs, int64_t R, const __nv_bfloat16* __restrict__ WD /*[128][320]*/,
                const float* __restrict__ f32sec, float* __restrict__ dirvec /*[R][128]*/) {
  __shared__ float s_w[27][128];                       // W_d[:, 256:283] transposed, fp32 copy of the bf16 values
  __shared__ float s_b[128];
  for (int i = threadIdx.x; i < 27 * 128; i += blockDim.x) {
    const int f = i / 128, j = i - f * 128;
    s_w[f][j] = __bfloat162float(WD[(size_t)j * 320 + 256 + f]);
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_b[i] = f32sec[kBD + i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const float d[3] = {__ldg(dirs + r * 3), __ldg(dirs + r * 3 + 1), __ldg(dirs + r * 3 + 2)};
    float feat[27];
    pe_features_fast<kDirFreqs>(d, feat);
    float acc[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] = s_b[lane + 32 * m];
#pragma unroll
    for (int f = 0; f < 27; ++f) {
      const float v = __bfloat162float(__float2bfloat16_rn(feat[f]));
#pragma unroll
      for (int m = 0; m < 4; ++m) acc[m] = fmaf(v, s_w[f][lane + 32 * m], acc[m]);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) dirvec[r * 128 + lane + 32 * m] = acc[m];
  }
}

// generic fp32 PositionalEncoding.forward / backward (API parity; C components per vector)
__global__ void posenc_fwd_kernel(const float* __restrict__ x, int64_t n, int C, int L, float* __restrict__ out) {
  const int W = C * (1 + 2 * L);
  const int64_t total = n * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int a = (int)(i - r * C);
    const float v = x[i];
    float* o = out + r * W;
    o[a] = v;
    for (int k = 0; k < L; ++k) {
      float s, c;
      sincosf(__fmul_rn((float)(1 << k), v), &s, &c);
      o[C * (1 + 2 * k) + a] = s;
      o[C * (2 + 2 * k) + a] = c;
    }
  }
}
__global__ void posenc_bwd_kernel(const float* __restrict__ x, int64_t n, int C, int L, const float* __restrict__ g,
                                  float* __restrict__ gx) {
  const int W = C * (1 + 2 * L);
  const int64_t total = n * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int a = (int)(i - r * C);
    const float v = x[i];
    const float* go = g + r * W;
    float acc = go[a];
    for (int k = 0; k < L; ++k) {
      const float f = (float)(1 << k);
      float s, c;
      sincosf(__fmul_rn(f, v), &s, &c);
      acc += f * (c * go[C * (1 + 2 * k) + a] - s * go[C * (2 + 2 * k) + a]);
    }
    gx[i] = acc;
  }
}

// backward of the fused encode: g_pts from dXE0 + dXE5 (bf16 [M,64]); g_dirs from dDE (bf16 [M,64])
template <int L>
__device__ __forceinline__ void encode3_bwd(const float x[3], const float* g /*3*(1+2L)*/, float gx[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) gx[a] = g[a];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, c;
      sincosf(__fmul_rn(f, x[a]), &s, &c);
      gx[a] += f * (c * g[3 + 6 * k + a] - s * g[3 + 6 * k + 3 + a]);
    }
  }
}

__global__ void __launch_bounds__(256)
encode_bwd_pts_kernel(const float* __restrict__ pts, int64_t M, const __nv_bfloat16* __restrict__ dXE0,
                      const __nv_bfloat16* __restrict__ dXE5, float* __restrict__ g_pts) {
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    float g[64];
    const uint4* a = reinterpret_cast<const uint4*>(dXE0 + m * 64);
    const uint4* b = reinterpret_cast<const uint4*>(dXE5 + m * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = a[i], v = b[i];
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        g[8 * i + 2 * e] = bf16_lo(uw[e]) + bf16_lo(vw[e]);
        g[8 * i + 2 * e + 1] = bf16_hi(uw[e]) + bf16_hi(vw[e]);
      }
    }
    const float x[3] = {pts[m * 3], pts[m * 3 + 1], pts[m * 3 + 2]};
    float gx[3];
    encode3_bwd<kPosFreqs>(x, g, gx);
    g_pts[m * 3] = gx[0]; g_pts[m * 3 + 1] = gx[1]; g_pts[m * 3 + 2] = gx[2];
  }
}

// one warp per direction group: g_dirs[r] = sum over the group's points of dPE^T dDE[m]
__global__ void __launch_bounds__(256)
encode_bwd_dirs_kernel(const float* __restrict__ dirs, int64_t R, int group, const __nv_bfloat16* __restrict__ dDE,
                       float* __restrict__ g_dirs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    float g[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) g[i] = 0.f;
    for (int p = lane; p < group; p += 32) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(dDE + (r * group + p) * 64);
#pragma unroll
      for (int i = 0; i < 13; ++i) { const uint32_t w = src[i]; g[2 * i] += bf16_lo(w); g[2 * i + 1] += bf16_hi(w); }
      g[26] += bf16_lo(src[13]);
    }
#pragma unroll
    for (int i = 0; i < 27; ++i) g[i] = warp_sum(g[i]);
    if (lane == 0) {
      const float d[3] = {dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]};
      float gx[3];
      encode3_bwd<kDirFreqs>(d, g, gx);
      g_dirs[r * 3] = gx[0]; g_dirs[r * 3 + 1] = gx[1]; g_dirs[r * 3 + 2] = gx[2];
    }
  }
}

// ------------------------------------------------------------------------------------------
// heads backward (the forward heads are fused into the GEMM epilogues, gemm_tcgen05.cu).
// thread <-> (point, 8-column group of HC): coalesced 16 B loads/stores, 16 threads per point.
//   dHC[m, j]      = (HC[m, j] > 0) * sum_c g_raw[m, c] * W_rgb[c, j]                (bf16)
//   dFS[m, 256..]  = (d sigma_pre, 0, ...)                                            (bf16)
//   dW_rgb[c, j]  += g_raw[m, c] * HC[m, j],  db_rgb[c] += g_raw[m, c]   per-thread partials, reduced
//                    in a fixed order: block tree here, heads_bwd_reduce_kernel across CTAs.
// ------------------------------------------------------------------------------------------
constexpr int kHeadsBwdThreads = 256;
constexpr int kHeadsBwdCtasPerSm = 4;
constexpr int kHeadsBwdPartial = 648;       // per CTA: dW_rgb 384 | db_rgb 3 | pad | dW_sigma 256 (at 388) | db_sigma (at 644) | pad

// SIGMA: also the sigma_linear gradients, dW_sigma[j] += dsigma[m] * H7[m, j] and db_sigma += dsigma[m] with dsigma rounded
// to bf16 exactly as the tensor-core path sees it (dFS[m, 256]) -- 16 columns of H7 per thread.  The weight-gradient
// stream (wgrad_stream.cu) would need a CTA pair per split for this one row of dFS^T; here it costs 256 B per point of
// extra reads in a kernel that is already streaming HC.
template <bool SIGMA>
__global__ void __launch_bounds__(kHeadsBwdThreads)
heads_bwd_kernel(const float4* __restrict__ g_raw, const __nv_bfloat16* __restrict__ HC, const __nv_bfloat16* __restrict__ H7,
                 int64_t M, const float* __restrict__ f32sec, __nv_bfloat16* __restrict__ dHC,
                 __nv_bfloat16* __restrict__ dFS, int ldfs, float* __restrict__ partial /*[grid][kHeadsBwdPartial]*/) {
  __shared__ float s_red[kHeadsBwdThreads / 16][16][25];
  const int cg = threadIdx.x & 15;            // column group: columns 8*cg .. 8*cg+7 of HC (16*cg .. 16*cg+15 of H7)
  const int pl = threadIdx.x >> 4;            // point lane inside the CTA
  float w[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) w[c][e] = f32sec[kWRgb + c * 128 + cg * 8 + e];
  float acc[3][8], bacc[3] = {0.f, 0.f, 0.f};
  float sacc[SIGMA ? 16 : 1], sbias = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
#pragma unroll
  for (int e = 0; e < (SIGMA ? 16 : 1); ++e) sacc[e] = 0.f;
  const int64_t pstride = (int64_t)gridDim.x * (kHeadsBwdThreads / 16);
  // several points per thread and iteration, all loads issued before any dependent math: one uint4 per thread in flight
  // (16 KB per SM) left this HBM-bound kernel at half of its bandwidth floor
  constexpr int kU = SIGMA ? 2 : 4;
  for (int64_t m0 = (int64_t)blockIdx.x * (kHeadsBwdThreads / 16) + pl; m0 < M; m0 += kU * pstride) {
    float4 gv[kU];
    uint4 hvv[kU];
    uint4 h7v[SIGMA ? kU : 1][2];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t m = m0 + u * pstride;
      if (m < M) {
        gv[u] = __ldg(g_raw + m);
        hvv[u] = __ldcs(reinterpret_cast<const uint4*>(HC + m * 128 + cg * 8));
        if (SIGMA) {
          h7v[u][0] = __ldcs(reinterpret_cast<const uint4*>(H7 + m * 256 + cg * 16));
          h7v[u][1] = __ldcs(reinterpret_cast<const uint4*>(H7 + m * 256 + cg * 16 + 8));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t m = m0 + u * pstride;
      if (m >= M) break;
      const float4 g = gv[u];
      const uint32_t hw[4] = {hvv[u].x, hvv[u].y, hvv[u].z, hvv[u].w};
      uint32_t outw[4];
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        const float h0 = bf16_lo(hw[e2]), h1 = bf16_hi(hw[e2]);
        acc[0][2 * e2] = fmaf(g.x, h0, acc[0][2 * e2]); acc[0][2 * e2 + 1] = fmaf(g.x, h1, acc[0][2 * e2 + 1]);
        acc[1][2 * e2] = fmaf(g.y, h0, acc[1][2 * e2]); acc[1][2 * e2 + 1] = fmaf(g.y, h1, acc[1][2 * e2 + 1]);
        acc[2][2 * e2] = fmaf(g.z, h0, acc[2][2 * e2]); acc[2][2 * e2 + 1] = fmaf(g.z, h1, acc[2][2 * e2 + 1]);
        const float d0 = h0 > 0.f ? (g.x * w[0][2 * e2] + g.y * w[1][2 * e2] + g.z * w[2][2 * e2]) : 0.f;
        const float d1 = h1 > 0.f ? (g.x * w[0][2 * e2 + 1] + g.y * w[1][2 * e2 + 1] + g.z * w[2][2 * e2 + 1]) : 0.f;
        outw[e2] = pack_bf16(d0, d1);
      }
      *reinterpret_cast<uint4*>(dHC + m * 128 + cg * 8) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
      const uint32_t ds_packed = pack_bf16(g.w, 0.f);
      if (SIGMA) {
        const float ds = bf16_lo(ds_packed);               // what dFS[m, 256] holds
        const uint32_t w7[8] = {h7v[u][0].x, h7v[u][0].y, h7v[u][0].z, h7v[u][0].w, h7v[u][1].x, h7v[u][1].y, h7v[u][1].z, h7v[u][1].w};
#pragma unroll
        for (int e2 = 0; e2 < 8; ++e2) {
          sacc[2 * e2] = fmaf(ds, bf16_lo(w7[e2]), sacc[2 * e2]);
          sacc[2 * e2 + 1] = fmaf(ds, bf16_hi(w7[e2]), sacc[2 * e2 + 1]);
        }
        if (cg == 0) sbias += ds;
      }
      if (cg == 0) {
        bacc[0] += g.x; bacc[1] += g.y; bacc[2] += g.z;
        *reinterpret_cast<uint4*>(dFS + m * ldfs + 256) = make_uint4(ds_packed, 0, 0, 0);
      } else if (cg == 1) {
        *reinterpret_cast<uint4*>(dFS + m * ldfs + 264) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  // fixed-order reduction over the 16 point lanes of the CTA
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) s_red[pl][cg][c * 8 + e] = acc[c][e];
  s_red[pl][cg][24] = 0.f;
  __syncthreads();
  float* p = partial + (size_t)blockIdx.x * kHeadsBwdPartial;
  for (int o = threadIdx.x; o < 384; o += kHeadsBwdThreads) {
    const int c = o / 128, j = o % 128;
    float a = 0.f;
    for (int l = 0; l < kHeadsBwdThreads / 16; ++l) a += s_red[l][j >> 3][c * 8 + (j & 7)];
    p[o] = a;
  }
  __syncthreads();
  if (cg == 0) { s_red[pl][0][0] = bacc[0]; s_red[pl][0][1] = bacc[1]; s_red[pl][0][2] = bacc[2]; s_red[pl][0][3] = sbias; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a = 0.f;
    for (int l = 0; l < kHeadsBwdThreads / 16; ++l) a += s_red[l][0][threadIdx.x];
    p[384 + threadIdx.x] = a;
  }
  if (SIGMA) {
    if (threadIdx.x == 3) {
      float a = 0.f;
      for (int l = 0; l < kHeadsBwdThreads / 16; ++l) a += s_red[l][0][3];
      p[644] = a;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; ++e) s_red[pl][cg][e] = sacc[e];
    __syncthreads();
    {
      const int j = threadIdx.x;               // 256 threads <-> 256 columns of H7
      float a = 0.f;
      for (int l = 0; l < kHeadsBwdThreads / 16; ++l) a += s_red[l][j >> 4][j & 15];
      p[388 + j] = a;
    }
  }
}
// one warp per output, lanes over the per-CTA partials, fixed shuffle tree.  Outputs 0..383 dW_rgb, 384..386 db_rgb and, when
// gWsig is given, 387..642 dW_sigma, 643 db_sigma.
__global__ void heads_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ gWrgb,
                                        float* __restrict__ gBrgb, float* __restrict__ gWsig, float* __restrict__ gBsig) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_out = gWsig ? 644 : 387;
  if (o >= n_out) return;
  const int src = o < 387 ? o : (o < 643 ? 388 + (o - 387) : 644);
  float acc = 0.f;
  for (int b = lane; b < nblk; b += 32) acc += partial[(size_t)b * kHeadsBwdPartial + src];
  acc = warp_sum(acc);
  if (lane == 0) {
    if (o < 384) gWrgb[o] = acc;
    else if (o < 387) gBrgb[o - 384] = acc;
    else if (o < 643) gWsig[o - 387] = acc;
    else gBsig[0] = acc;
  }
}

// model.py:181,194 activations for the NeRF.forward API
__global__ void head_act_fwd_kernel(const float4* __restrict__ raw, int64_t M, float* __restrict__ rgb, float* __restrict__ sigma) {
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float4 r = raw[m];
    rgb[m * 3] = 1.0f / (1.0f + expf(-r.x)); rgb[m * 3 + 1] = 1.0f / (1.0f + expf(-r.y)); rgb[m * 3 + 2] = 1.0f / (1.0f + expf(-r.z));
    sigma[m] = fmaxf(r.w, 0.f);
  }
}
__global__ void head_act_bwd_kernel(const float4* __restrict__ raw, int64_t M, const float* __restrict__ g_rgb,
                                    const float* __restrict__ g_sigma, float4* __restrict__ g_raw) {
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float4 r = raw[m];
    const float s0 = 1.0f / (1.0f + expf(-r.x)), s1 = 1.0f / (1.0f + expf(-r.y)), s2 = 1.0f / (1.0f + expf(-r.z));
    float4 o;
    o.x = g_rgb ? g_rgb[m * 3] * s0 * (1.f - s0) : 0.f;
    o.y = g_rgb ? g_rgb[m * 3 + 1] * s1 * (1.f - s1) : 0.f;
    o.z = g_rgb ? g_rgb[m * 3 + 2] * s2 * (1.f - s2) : 0.f;
    o.w = (g_sigma && r.w > 0.f) ? g_sigma[m] : 0.f;
    g_raw[m] = o;
  }
}

// ------------------------------------------------------------------------------------------
// weight packing: 24 fp32 state_dict tensors -> padded bf16 matrices + fp32 section
// ------------------------------------------------------------------------------------------
struct PackSrc { const float* p[RN_NUM_PARAM_TENSORS]; };

struct PackSeg { size_t dst; int rows_p, cols_p, src_rows, src_cols, src_idx, skip; };

__constant__ PackSeg c_segs[11] = {
    {kW0, 256, 64, 256, 63, 0, 0},   {kW1, 256, 256, 256, 256, 2, 0}, {kW2, 256, 256, 256, 256, 4, 0},
    {kW3, 256, 256, 256, 256, 6, 0}, {kW4, 256, 256, 256, 256, 8, 0}, {kW5, 256, 320, 256, 319, 10, 1},
    {kW6, 256, 256, 256, 256, 12, 0}, {kW7, 256, 256, 256, 256, 14, 0},
    {kWFS, 256, 256, 256, 256, 18, 0},            // feature_linear.weight -> rows 0..255 of WFS
    {kWFS + 256 * 256, 16, 256, 1, 256, 16, 0},   // sigma_linear.weight -> row 256 (+15 zero rows)
    {kWD, 128, 320, 128, 283, 20, 0}};

__global__ void pack_weights_kernel(PackSrc src, __nv_bfloat16* __restrict__ dst_bf16, float* __restrict__ dst_f32) {
  const int seg = blockIdx.y;
  if (seg < 11) {
    const PackSeg sg = c_segs[seg];
    const int total = sg.rows_p * sg.cols_p;
    const float* s = src.p[sg.src_idx];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int r = i / sg.cols_p, c = i % sg.cols_p;
      int sc = c;
      if (sg.skip) sc = (c < 63) ? c : (c == 63 ? -1 : c - 1);   // [x_enc(63) | 0 | h(256)]
      float v = 0.f;
      if (r < sg.src_rows && sc >= 0 && sc < sg.src_cols) v = s[(size_t)r * sg.src_cols + sc];
      dst_bf16[sg.dst + i] = __float2bfloat16_rn(v);
    }
  } else {
    // fp32 section
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)kF32Elems; i += gridDim.x * blockDim.x) {
      float v = 0.f;
      if (i < (int)kBF) v = src.p[2 * (i / 256) + 1][i % 256];                 // trunk biases
      else if (i < (int)kBD) v = src.p[19][i - kBF];                            // feature bias
      else if (i < (int)kWSig) v = src.p[21][i - kBD];                          // dir bias
      else if (i < (int)kBSig) v = src.p[16][i - kWSig];                        // sigma weight
      else if (i < (int)kWRgb) v = (i == (int)kBSig) ? src.p[17][0] : 0.f;      // sigma bias
      else if (i < (int)kBRgb) v = src.p[22][i - kWRgb];                        // rgb weight
      else v = (i - (int)kBRgb < 3) ? src.p[23][i - kBRgb] : 0.f;               // rgb bias
      dst_f32[i] = v;
    }
  }
}

}  // namespace rn

using namespace rn;
using namespace rn::layout;

namespace rn {
// used by mlp.cu
int launch_encode(const float* pts, const float* dirs, int64_t M, int group, void* XC, int ldx, void* FD, int ldf,
                  cudaStream_t st) {
  encode_kernel<<<grid_for(M, 256), 256, 0, st>>>(pts, dirs, M, group, (__nv_bfloat16*)XC, ldx, (__nv_bfloat16*)FD, ldf);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
int launch_dir_bias(const float* dirs, int64_t R, const void* packed, float* dirvec, cudaStream_t st) {
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(packed);
  const float* F = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + kBf16Bytes);
  dir_bias_kernel<<<grid_for(R * 32, 256, 4), 256, 0, st>>>(dirs, R, W + kWD, F, dirvec);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
size_t heads_bwd_scratch_bytes(int64_t M) { return (size_t)num_sms() * kHeadsBwdCtasPerSm * kHeadsBwdPartial * sizeof(float) + 256; }
// H7 != null: also the sigma_linear gradients (gWsig[256], gBsig[1]) -- the path taken when the weight-gradient stream runs
int launch_heads_bwd(const float* g_raw, const void* HC, int64_t M, const float* f32sec, void* dHC, void* dFS, int ldfs,
                     float* scratch, float* gWrgb, float* gBrgb, cudaStream_t st, const void* H7, float* gWsig, float* gBsig) {
  const int nblk = num_sms() * kHeadsBwdCtasPerSm;
  if (H7) {
    RN_REQUIRE(gWsig && gBsig);
    heads_bwd_kernel<true><<<nblk, kHeadsBwdThreads, 0, st>>>((const float4*)g_raw, (const __nv_bfloat16*)HC,
                                                              (const __nv_bfloat16*)H7, M, f32sec, (__nv_bfloat16*)dHC,
                                                              (__nv_bfloat16*)dFS, ldfs, scratch);
  } else {
    heads_bwd_kernel<false><<<nblk, kHeadsBwdThreads, 0, st>>>((const float4*)g_raw, (const __nv_bfloat16*)HC, nullptr, M, f32sec,
                                                               (__nv_bfloat16*)dHC, (__nv_bfloat16*)dFS, ldfs, scratch);
  }
  RN_LAUNCH_CHECK();
  const int n_out = H7 ? 644 : 387;
  heads_bwd_reduce_kernel<<<(n_out * 32 + 255) / 256, 256, 0, st>>>(scratch, nblk, gWrgb, gBrgb, H7 ? gWsig : nullptr, gBsig);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
int launch_encode_bwd(const float* pts, const float* dirs, int64_t M, int group, const void* dXE0, const void* dXE5,
                      const void* dDE, float* g_pts, float* g_dirs, cudaStream_t st) {
  if (g_pts) {
    encode_bwd_pts_kernel<<<grid_for(M, 256), 256, 0, st>>>(pts, M, (const __nv_bfloat16*)dXE0, (const __nv_bfloat16*)dXE5, g_pts);
    RN_LAUNCH_CHECK();
  }
  if (g_dirs) {
    const int64_t R = M / group;
    encode_bwd_dirs_kernel<<<grid_for(R * 32, 256), 256, 0, st>>>(dirs, R, group, (const __nv_bfloat16*)dDE, g_dirs);
    RN_LAUNCH_CHECK();
  }
  return RN_OK;
}
}  // namespace rn

extern "C" {

size_t rn_mlp_packed_weight_bytes(void) { return kPackedBytes; }

int rn_mlp_pack_weights(const float* const* params_host, void* packed, rn_stream_t stream) {
  RN_REQUIRE(params_host && packed);
  PackSrc src;
  for (int i = 0; i < RN_NUM_PARAM_TENSORS; ++i) {
    RN_REQUIRE(params_host[i] != nullptr);
    src.p[i] = params_host[i];
  }
  pack_weights_kernel<<<dim3(64, 12), 256, 0, (cudaStream_t)stream>>>(
      src, (__nv_bfloat16*)packed, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + kBf16Bytes));
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_head_act_fwd(const float* raw4, int64_t M, float* rgb, float* sigma, rn_stream_t stream) {
  if (M == 0) return RN_OK;
  RN_REQUIRE(raw4 && rgb && sigma && M >= 0);
  head_act_fwd_kernel<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)raw4, M, rgb, sigma);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_head_act_bwd(const float* raw4, int64_t M, const float* g_rgb, const float* g_sigma, float* g_raw4, rn_stream_t stream) {
  if (M == 0) return RN_OK;
  RN_REQUIRE(raw4 && g_raw4 && M >= 0);
  head_act_bwd_kernel<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)raw4, M, g_rgb, g_sigma, (float4*)g_raw4);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_posenc_fwd(const float* x, int64_t n, int C, int L, float* out, rn_stream_t stream) {
  if (n == 0) return RN_OK;
  RN_REQUIRE(x && out && n >= 0 && C >= 1 && L >= 0 && L <= 30);
  posenc_fwd_kernel<<<grid_for(n * C, 256), 256, 0, (cudaStream_t)stream>>>(x, n, C, L, out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_posenc_bwd(const float* x, int64_t n, int C, int L, const float* g_out, float* g_x, rn_stream_t stream) {
  if (n == 0) return RN_OK;
  RN_REQUIRE(x && g_out && g_x && n >= 0 && C >= 1 && L >= 0 && L <= 30);
  posenc_bwd_kernel<<<grid_for(n * C, 256), 256, 0, (cudaStream_t)stream>>>(x, n, C, L, g_out, g_x);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // extern "C"
