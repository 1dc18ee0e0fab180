// sampling.cu -- stratified sampling, inverse-CDF (hierarchical) resampling, point generation.
//
// Replaces noisy_src/rays.py:145-210 (sample_along_rays), :213-279 (sample_pdf) and
// :282-333 (sample_hierarchical): ~35 aten launches (linspace/cat/rand-mul-add, sum, div, cumsum,
// searchsorted, clamp, gather x2, where, sort, broadcast mul-add) become one launch each.
//
// sample_pdf / sample_hierarchical are warp-per-ray kernels: the ray's bins, cdf and merge buffer
// live in shared memory, each lane inverts the cdf for Nf/32 draws with a binary search
// (searchsorted right=True semantics), and the 32 lanes sort the merged depths with a bitonic
// network.  The cdf is a warp scan in a fixed order that oracle/nerf_oracle.py restates, so
// sample indices are bit-exact against the oracle; all interpolation arithmetic is unfused
// (__fmul_rn/__fadd_rn) to reproduce the reference's separate mul/add roundings.
//
// HBM-bound: 20*Nc+24 B/ray (stratified), 24*Nc+20*Nf+24 B/ray (hierarchical incl. pts).
#include "common.cuh"
#include <math_constants.h>

namespace rn {

// ------------------------------------------------------------------------------------------
// stratified: z = lower + (upper - lower) * t_rand ; pts = o + d*z         (rays.py:197-208)
// ------------------------------------------------------------------------------------------
__global__ void stratified_kernel(const float* __restrict__ ro, const float* __restrict__ rd, int64_t B,
                                  const float* __restrict__ zb, int Nc, const float* __restrict__ t_rand,
                                  float* __restrict__ z_out, float* __restrict__ pts) {
  const int64_t n = B * Nc;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // four samples per thread and iteration, their loads issued first (the kernel is latency-bound:
  // profiles/r01_ncu_full_hbm_kernels.raw.csv, 20 long-scoreboard stall cycles per issued instruction)
  constexpr int kU = 4;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += kU * stride) {
    float tr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      tr[u] = (t_rand && i < n) ? __ldcs(t_rand + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n) break;
      const int64_t b = i / Nc;
      const int s = (int)(i - b * Nc);
      float z = __ldg(zb + s);
      if (t_rand) {
        const float lower = (s == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, __ldg(zb + s - 1)));
        const float upper = (s == Nc - 1) ? z : __fmul_rn(0.5f, __fadd_rn(__ldg(zb + s + 1), z));
        z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr[u]));
      }
      z_out[i] = z;
      if (pts) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
          pts[i * 3 + k] = __fadd_rn(__ldg(ro + b * 3 + k), __fmul_rn(__ldg(rd + b * 3 + k), z));
      }
    }
  }
}

// Four consecutive samples of one ray per thread (Nc % 4 == 0, 16-byte aligned buffers): one 16-byte load of the draws,
// one 16-byte store of depths, three of points -- full 32-byte sectors per instruction instead of three 4-byte stores at a
// 12-byte stride (the scalar kernel reached 0.41 of the HBM roofline at configs[4], profiles/r02_ab_log.md).
__global__ void __launch_bounds__(256)
stratified_vec4_kernel(const float* __restrict__ ro, const float* __restrict__ rd, int64_t B, const float* __restrict__ zb,
                       int Nc, const float4* __restrict__ t_rand, float4* __restrict__ z_out, float4* __restrict__ pts) {
  const int q_per_ray = Nc >> 2;
  const int64_t n = B * q_per_ray;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t b = i / q_per_ray;
    const int s0 = (int)(i - b * q_per_ray) << 2;
    float4 tr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t_rand) tr = __ldcs(t_rand + i);
    const float trv[4] = {tr.x, tr.y, tr.z, tr.w};
    // base depths s0-1 .. s0+4 (clamped at the ends: the end bins use the end depth itself, rays.py:199-201)
    float zz[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) zz[k] = __ldg(zb + min(max(s0 - 1 + k, 0), Nc - 1));
    float z[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int s = s0 + k;
      z[k] = zz[k + 1];
      if (t_rand) {
        const float lower = (s == 0) ? zz[k + 1] : __fmul_rn(0.5f, __fadd_rn(zz[k + 1], zz[k]));
        const float upper = (s == Nc - 1) ? zz[k + 1] : __fmul_rn(0.5f, __fadd_rn(zz[k + 2], zz[k + 1]));
        z[k] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), trv[k]));
      }
    }
    __stcs(z_out + i, make_float4(z[0], z[1], z[2], z[3]));
    if (pts) {
      const float o0 = __ldg(ro + b * 3), o1 = __ldg(ro + b * 3 + 1), o2 = __ldg(ro + b * 3 + 2);
      const float d0 = __ldg(rd + b * 3), d1 = __ldg(rd + b * 3 + 1), d2 = __ldg(rd + b * 3 + 2);
      float4* pq = pts + 3 * i;
      __stcs(pq, make_float4(__fadd_rn(o0, __fmul_rn(d0, z[0])), __fadd_rn(o1, __fmul_rn(d1, z[0])),
                             __fadd_rn(o2, __fmul_rn(d2, z[0])), __fadd_rn(o0, __fmul_rn(d0, z[1]))));
      __stcs(pq + 1, make_float4(__fadd_rn(o1, __fmul_rn(d1, z[1])), __fadd_rn(o2, __fmul_rn(d2, z[1])),
                                 __fadd_rn(o0, __fmul_rn(d0, z[2])), __fadd_rn(o1, __fmul_rn(d1, z[2]))));
      __stcs(pq + 2, make_float4(__fadd_rn(o2, __fmul_rn(d2, z[2])), __fadd_rn(o0, __fmul_rn(d0, z[3])),
                                 __fadd_rn(o1, __fmul_rn(d1, z[3])), __fadd_rn(o2, __fmul_rn(d2, z[3]))));
    }
  }
}

__global__ void points_fwd_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                                  const float* __restrict__ z, int64_t B, int S, float* __restrict__ pts) {
  const int64_t n = B * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / S;
    const float zz = z[i];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      pts[i * 3 + k] = __fadd_rn(__ldg(ro + b * 3 + k), __fmul_rn(__ldg(rd + b * 3 + k), zz));
  }
}

// g_o[b] = sum_s g_pts[b,s,:], g_d[b] = sum_s z[b,s] * g_pts[b,s,:]   (warp per ray)
__global__ void points_bwd_kernel(const float* __restrict__ g_pts, const float* __restrict__ z, int64_t B, int S,
                                  float* __restrict__ g_o, float* __restrict__ g_d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    float ao[3] = {0.f, 0.f, 0.f}, ad[3] = {0.f, 0.f, 0.f};
    for (int s = lane; s < S; s += 32) {
      const float zz = z[b * S + s];
      const float* g = g_pts + (b * S + s) * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float v = g[k]; ao[k] += v; ad[k] = fmaf(zz, v, ad[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { ao[k] = warp_sum(ao[k]); ad[k] = warp_sum(ad[k]); }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { g_o[b * 3 + k] = ao[k]; g_d[b * 3 + k] = ad[k]; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// warp-level pieces of sample_pdf
// ------------------------------------------------------------------------------------------
// inclusive Kogge-Stone scan of one value per lane (the order oracle/nerf_oracle.py:_warp_scan restates)
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = __fadd_rn(v, n);
  }
  return v;
}

// s_w[0..nb-2] holds raw weights on entry; on exit s_cdf[0..nb-1] holds the cdf (s_cdf[0]=0).
// Warp-scan cdf: lane l owns the contiguous chunk [l*C, (l+1)*C) (C = ceil(nw / 32)), sums it left to right, the lane
// totals go through a Kogge-Stone scan, element k gets offset[lane] + local prefix.  The oracle sums in exactly this
// order, so the searchsorted indices stay bit-exact; torch's own order is backend-dependent (pairwise sum + sequential
// cumsum on the CPU, block scan on CUDA) and differs from any fixed order in <= 2e-3 of the draws
// (tests/test_oracle_golden.py).  One lane adding all nb values sequentially was a third of the kernel's instructions.
__device__ __forceinline__ void warp_build_cdf(float* s_w, float* s_cdf, int nb, int lane) {
  const int nw = nb - 1;
  const int C = (nw + 31) >> 5;
  const int k0 = lane * C;
  float run = 0.f;
  for (int j = 0; j < C; ++j) {
    const int k = k0 + j;
    if (k < nw) {
      const float x = __fadd_rn(s_w[k], 1e-5f);                               // rays.py:243
      s_w[k] = x;
      run = __fadd_rn(run, x);
    }
  }
  const float tot = __shfl_sync(0xffffffffu, warp_scan_add(run, lane), 31);   // rays.py:246 denominator
  run = 0.f;
  for (int j = 0; j < C; ++j) {
    const int k = k0 + j;
    if (k < nw) {
      run = __fadd_rn(run, __fdiv_rn(s_w[k], tot));                           // pdf, local prefix
      s_w[k] = run;
    }
  }
  const float incl = warp_scan_add(run, lane);
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) { excl = 0.f; s_cdf[0] = 0.f; }
  for (int j = 0; j < C; ++j) {
    const int k = k0 + j;
    if (k < nw) s_cdf[k + 1] = __fadd_rn(excl, s_w[k]);                       // rays.py:247-248
  }
  __syncwarp();
}

// one draw: searchsorted(right=True) + clamp + gather + interpolation (rays.py:259-277)
__device__ __forceinline__ float invert_cdf(const float* s_cdf, const float* s_bins, int nb, float u, int& ind) {
  int lo = 0, hi = nb;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  ind = lo;
  const int below = max(lo - 1, 0), above = min(lo, nb - 1);
  const float cb = s_cdf[below], ca = s_cdf[above], bb = s_bins[below], ba = s_bins[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.0f;
  const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

// bitonic sort of s[0..P) (P power of two) by one warp, ascending
__device__ __forceinline__ void warp_bitonic_sort(float* s, int P, int lane) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        // t-th compare-exchange pair of this stage
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const bool up = ((i & k) == 0);
        const float a = s[i], b = s[p];
        if ((a > b) == up) { s[i] = b; s[p] = a; }
      }
      __syncwarp();
    }
  }
}

constexpr int kWarpsPerCta = 8;
// floats of shared memory per warp of sample_hierarchical_kernel: z, bins, w, cdf [Nc each]; u, samples [NS each];
// merged depths [Nc + Nf]; bucket starts [NS + 1] ints; two uint16 index arrays [NS each] (= NS floats); + padding
__host__ __device__ constexpr int hier_smem_floats(int Nc, int NS, int Nf, bool inds) {
  return 4 * Nc + 2 * NS + (Nc + Nf) + (NS + 1) + (inds ? NS : 0) + 3;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, int64_t B, int nb,
                  const float* __restrict__ u, int64_t u_stride, int Nf, float* __restrict__ samples,
                  int64_t* __restrict__ inds_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* s_bins = smem + (size_t)w * (3 * nb);
  float* s_w = s_bins + nb;
  float* s_cdf = s_w + nb;
  for (int64_t b = blockIdx.x * (int64_t)kWarpsPerCta + w; b < B; b += (int64_t)gridDim.x * kWarpsPerCta) {
    for (int k = lane; k < nb; k += 32) s_bins[k] = bins[b * nb + k];
    for (int k = lane; k < nb - 1; k += 32) s_w[k] = weights[b * (nb - 1) + k];
    __syncwarp();
    warp_build_cdf(s_w, s_cdf, nb, lane);
    for (int j = lane; j < Nf; j += 32) {
      int ind;
      const float v = invert_cdf(s_cdf, s_bins, nb, u[b * u_stride + j], ind);
      samples[b * Nf + j] = v;
      if (inds_out) inds_out[b * Nf + j] = ind;
    }
    __syncwarp();
  }
}

// ---- register-resident bitonic sort: 32*E values, element index i = lane*E + e ----
template <int E>
__device__ __forceinline__ void warp_bitonic_sort_regs(float (&v)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= E) {
        const int lj = j / E;                       // partner lane distance
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float other = __shfl_xor_sync(0xffffffffu, v[e], lj);
          const int i = lane * E + e;
          const bool up = (i & k) == 0;
          const bool lower = (i & j) == 0;
          v[e] = (lower == up) ? fminf(v[e], other) : fmaxf(v[e], other);
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & j) == 0) {
            const int i = lane * E + e;
            const bool up = (i & k) == 0;
            const float a = v[e], b = v[e | j];
            const float mn = fminf(a, b), mx = fmaxf(a, b);
            v[e] = up ? mn : mx;
            v[e | j] = up ? mx : mn;
          }
        }
      }
    }
  }
}

// Merge-path split of two ascending shared-memory arrays: how many elements of A are among the first d outputs of the merge
// (A goes first on ties).  Every lane of a warp takes an equal slice of the merged sequence, so the merge loops below run
// the same number of iterations in all lanes (the per-element binary searches / stepping loops they replace ran as long as
// the unluckiest lane: 25 % of the kernel's instructions and its two largest stall sites, profiles/r02_ncu_hier_v3.md).
__device__ __forceinline__ int merge_path_split(const float* A, int na, const float* Bm, int nbm, int d) {
  int lo = max(0, d - nbm), hi = min(d, na);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (A[mid] <= Bm[d - mid - 1]) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ascending sort of E registers (odd-even transposition network, E <= 8)
template <int E>
__device__ __forceinline__ void sort_regs(float (&v)[E]) {
#pragma unroll
  for (int pass = 0; pass < E; ++pass) {
#pragma unroll
    for (int i = (pass & 1); i + 1 < E; i += 2) {
      const float a = v[i], b = v[i + 1];
      v[i] = fminf(a, b);
      v[i + 1] = fmaxf(a, b);
    }
  }
}

// sample_hierarchical, fast path (Nf <= 32*E <= 256).  Work per ray is kept near the byte count instead of the
// 2 x 256 binary searches + 36-stage bitonic network of the first version (0.18 of the HBM roofline at 128 + 256 samples,
// 69 % issue-bound: profiles/r01_ncu_full_hbm_kernels.raw.csv):
//   1. the draws u are put in ascending order FIRST: deterministic draws (linspace) already are -- one vote; random draws
//      are uniform by construction (torch.rand), so a counting sort into 32 E buckets by floor(u * 32 E) leaves ~1 draw per
//      bucket; what is left to order inside the buckets is done by two passes of an E-register sorting network over
//      lane-contiguous windows (offset 0 and E / 2), and a vote -- any other distribution of caller-supplied draws
//      falls through to the register bitonic network and stays correct, only slower;
//   2. with ascending u the searchsorted indices of ALL draws come out of one merge of the cdf knots with the draws
//      (knots first on ties = the predicate #{cdf <= u} of torch.searchsorted(right=True)), split evenly over the lanes
//      by merge-path partitioning; the interpolation then runs on E consecutive draws per lane;
//   3. the samples come out ascending except for 1-ulp inversions at bin borders (fl(b + t (a - b)) may exceed a):
//      checked with a vote, and only then sorted (register bitonic network, the first version's hot loop);
//   4. the final order is a second merge-path merge, of the coarse depths with the ascending samples.
// Values are those of sort(cat(z, samples)) bit for bit.  INDS (tests): also returns the searchsorted index of every draw
// in the caller's order; that variant ranks inside the buckets exactly (value, slot) so the permutation is known.
template <int E, bool INDS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4)
sample_hierarchical_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                           const float* __restrict__ zc, const float* __restrict__ weights, int64_t B, int Nc,
                           const float* __restrict__ u, int64_t u_stride, int Nf,
                           float* __restrict__ z_all, float* __restrict__ pts, int64_t* __restrict__ inds_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nb = Nc - 1, Nt = Nc + Nf;
  constexpr int NS = 32 * E;                 // padded draw count = bucket count
  float* s_z = smem + (size_t)w * hier_smem_floats(Nc, NS, Nf, INDS);
  float* s_bins = s_z + Nc;
  float* s_w = s_bins + Nc;
  float* s_cdf = s_w + Nc;
  float* s_u = s_cdf + Nc;                   // [NS] draws (caller's order, then ascending)
  float* s_smp = s_u + NS;                   // [NS] samples, ascending (scratch of the counting sort before that)
  float* s_out = s_smp + NS;                 // [Nc + Nf] merged depths
  int* s_cnt = reinterpret_cast<int*>(s_out + Nt);                       // [NS + 1] bucket counts -> starts
  unsigned short* s_id = reinterpret_cast<unsigned short*>(s_cnt + NS + 1);    // INDS: [NS] original index of the sorted draw
  unsigned short* s_id2 = s_id + NS;                                     // INDS: [NS] scratch
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(z_all) | reinterpret_cast<uintptr_t>(pts) |
                        (uintptr_t)__cvta_generic_to_shared(s_out)) & 15) == 0;
  for (int64_t b = blockIdx.x * (int64_t)kWarpsPerCta + w; b < B; b += (int64_t)gridDim.x * kWarpsPerCta) {
    const float* zrow_in = zc + b * Nc;
    const float* wrow_in = weights + b * Nc + 1;                        // interior weights, rays.py:321
    const float* urow = u + b * u_stride;
    for (int k = lane; k < Nc; k += 32) s_z[k] = __ldcs(zrow_in + k);
    for (int k = lane; k < Nc - 2; k += 32) s_w[k] = __ldcs(wrow_in + k);
    float ur[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int j = e * 32 + lane;            // coalesced over u
      ur[e] = (j < Nf) ? __ldg(urow + j) : CUDART_INF_F;
      s_u[j] = ur[e];
      if (INDS) s_id[j] = (unsigned short)j;
    }
    __syncwarp();
    for (int k = lane; k < nb; k += 32) s_bins[k] = __fmul_rn(0.5f, __fadd_rn(s_z[k + 1], s_z[k]));   // rays.py:316
    // ---- 1. ascending draws ----
    bool ok = true;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int j = e * 32 + lane;
      if (j + 1 < Nf) ok = ok && (ur[e] <= s_u[j + 1]);
    }
    float uq[E];                              // ascending draws, lane-contiguous: uq[e] <-> position lane * E + e
    if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
      for (int e = 0; e < E; ++e) uq[e] = s_u[lane * E + e];
    } else {
      for (int k = lane; k <= NS; k += 32) s_cnt[k] = 0;
      __syncwarp();
      int bk[E], arr[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = e * 32 + lane;
        const float x = ur[e];
        int k = (x >= 1.0f) ? NS - 1 : ((x > 0.0f) ? (int)(x * (float)NS) : 0);
        bk[e] = k < NS - 1 ? k : NS - 1;
        arr[e] = (j < Nf) ? atomicAdd(&s_cnt[bk[e]], 1) : 0;
      }
      __syncwarp();
      // exclusive scan of the NS bucket counts: lane l owns buckets [l*E, (l+1)*E)
      int cnt[E], tot = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) { cnt[e] = s_cnt[lane * E + e]; tot += cnt[e]; }
      int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
      int run = incl - tot;
      __syncwarp();
#pragma unroll
      for (int e = 0; e < E; ++e) { s_cnt[lane * E + e] = run; run += cnt[e]; }
      if (lane == 31) s_cnt[NS] = run;
#pragma unroll
      for (int e = 0; e < E; ++e) s_smp[e * 32 + lane] = CUDART_INF_F;      // padding beyond Nf stays +inf
      __syncwarp();
      // scatter into bucket order (arrival order inside a bucket)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = e * 32 + lane;
        if (j < Nf) {
          const int pos = s_cnt[bk[e]] + arr[e];
          s_smp[pos] = ur[e];
          if (INDS) s_id2[pos] = (unsigned short)j;
        }
      }
      __syncwarp();
      if (INDS) {
        // exact rank inside the bucket by (value, slot): the permutation is needed for inds_out
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int j = e * 32 + lane;
          if (j < Nf) {
            const int b0 = s_cnt[bk[e]], b1 = s_cnt[bk[e] + 1], pos = b0 + arr[e];
            const float x = ur[e];
            int rank = 0;
            for (int q = b0; q < b1; ++q) { const float y = s_smp[q]; rank += (y < x || (y == x && q < pos)) ? 1 : 0; }
            s_u[b0 + rank] = x;
            s_id[b0 + rank] = (unsigned short)j;
          }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) uq[e] = s_u[lane * E + e];
      } else {
        // two passes of an E-register sorting network over lane-contiguous windows (offsets 0 and E / 2): orders every
        // bucket that fits a window; anything else is caught by the vote below
#pragma unroll
        for (int e = 0; e < E; ++e) uq[e] = s_smp[lane * E + e];
        sort_regs<E>(uq);
        if (E >= 2) {
#pragma unroll
          for (int e = 0; e < E; ++e) s_smp[lane * E + e] = uq[e];
          __syncwarp();
          if (lane < 31) {
#pragma unroll
            for (int e = 0; e < E; ++e) uq[e] = s_smp[lane * E + E / 2 + e];
            sort_regs<E>(uq);
#pragma unroll
            for (int e = 0; e < E; ++e) s_smp[lane * E + E / 2 + e] = uq[e];
          }
          __syncwarp();
#pragma unroll
          for (int e = 0; e < E; ++e) uq[e] = s_smp[lane * E + e];
        }
        bool asc = true;
#pragma unroll
        for (int e = 0; e + 1 < E; ++e) asc = asc && (uq[e] <= uq[e + 1]);
        const float nxt = __shfl_down_sync(0xffffffffu, uq[0], 1);
        if (lane < 31) asc = asc && (uq[E - 1] <= nxt);
        if (!__all_sync(0xffffffffu, asc)) warp_bitonic_sort_regs<E>(uq, lane);
      }
    }
    __syncwarp();
    warp_build_cdf(s_w, s_cdf, nb, lane);
    // ---- 2. inversion: ind_j = #{k : cdf[k] <= u_j} for all draws by ONE merge of the cdf knots with the ascending draws ----
#pragma unroll
    for (int e = 0; e < E; ++e) s_u[lane * E + e] = uq[e];
    __syncwarp();
    {
      const int total = nb + Nf, per = (total + 31) >> 5;
      const int d0 = min(lane * per, total), d1 = min(d0 + per, total);
      int ai = merge_path_split(s_cdf, nb, s_u, Nf, d0), bi = d0 - ai;
      float ah = ai < nb ? s_cdf[ai] : CUDART_INF_F, bh = bi < Nf ? s_u[bi] : CUDART_INF_F;
      for (int d = d0; d < d1; ++d) {
        if (bi >= Nf || (ai < nb && ah <= bh)) {
          ++ai;
          ah = ai < nb ? s_cdf[ai] : CUDART_INF_F;
        } else {
          s_cnt[bi] = ai;                      // searchsorted(right=True) index of draw bi (bucket starts are dead by now)
          ++bi;
          bh = bi < Nf ? s_u[bi] : CUDART_INF_F;
        }
      }
    }
    __syncwarp();
    float v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int j = lane * E + e;
      v[e] = CUDART_INF_F;
      if (j < Nf) {
        const float x = uq[e];
        const int ind = s_cnt[j];
        const int below = max(ind - 1, 0), above = min(ind, nb - 1);
        const float cb = s_cdf[below], ca = s_cdf[above], bb = s_bins[below], ba = s_bins[above];
        float denom = __fsub_rn(ca, cb);
        if (denom < 1e-5f) denom = 1.0f;
        const float t = __fdiv_rn(__fsub_rn(x, cb), denom);
        v[e] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));                 // rays.py:259-277
        if (INDS) inds_out[b * Nf + s_id[j]] = ind;
      }
    }
    // ---- 4. ascending?  (1-ulp inversions at bin borders are possible; sort only then) ----
    {
      bool asc = true;
#pragma unroll
      for (int e = 0; e + 1 < E; ++e) asc = asc && (v[e] <= v[e + 1]);
      const float nxt = __shfl_down_sync(0xffffffffu, v[0], 1);
      if (lane < 31) asc = asc && (v[E - 1] <= nxt);
      if (!__all_sync(0xffffffffu, asc)) warp_bitonic_sort_regs<E>(v, lane);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) s_smp[lane * E + e] = v[e];
    __syncwarp();
    // ---- 3. merge of the coarse depths with the ascending samples (rays.py:328: sort of the concatenation, values only) ----
    {
      const int per = (Nt + 31) >> 5;
      const int d0 = min(lane * per, Nt), d1 = min(d0 + per, Nt);
      int ai = merge_path_split(s_z, Nc, s_smp, Nf, d0), bi = d0 - ai;
      float ah = ai < Nc ? s_z[ai] : CUDART_INF_F, bh = bi < Nf ? s_smp[bi] : CUDART_INF_F;
      for (int d = d0; d < d1; ++d) {
        if (bi >= Nf || (ai < Nc && ah <= bh)) {
          s_out[d] = ah;
          ++ai;
          ah = ai < Nc ? s_z[ai] : CUDART_INF_F;
        } else {
          s_out[d] = bh;
          ++bi;
          bh = bi < Nf ? s_smp[bi] : CUDART_INF_F;
        }
      }
    }
    __syncwarp();
    float* zrow = z_all + b * Nt;
    float* prow = pts ? pts + b * Nt * 3 : nullptr;
    if ((Nt & 3) == 0 && vec_ok) {
      // four depths per lane and step: one 16-byte store of depths, three of points (12 consecutive floats)
      const float o0 = pts ? ro[b * 3] : 0.f, o1 = pts ? ro[b * 3 + 1] : 0.f, o2 = pts ? ro[b * 3 + 2] : 0.f;
      const float d0 = pts ? rd[b * 3] : 0.f, d1 = pts ? rd[b * 3 + 1] : 0.f, d2 = pts ? rd[b * 3 + 2] : 0.f;
      for (int g = lane; g < (Nt >> 2); g += 32) {
        const float4 zq = *reinterpret_cast<const float4*>(s_out + 4 * g);
        __stcs(reinterpret_cast<float4*>(zrow) + g, zq);
        if (pts) {
          float4* pq = reinterpret_cast<float4*>(prow) + 3 * g;
          __stcs(pq, make_float4(__fadd_rn(o0, __fmul_rn(d0, zq.x)), __fadd_rn(o1, __fmul_rn(d1, zq.x)),
                                 __fadd_rn(o2, __fmul_rn(d2, zq.x)), __fadd_rn(o0, __fmul_rn(d0, zq.y))));
          __stcs(pq + 1, make_float4(__fadd_rn(o1, __fmul_rn(d1, zq.y)), __fadd_rn(o2, __fmul_rn(d2, zq.y)),
                                     __fadd_rn(o0, __fmul_rn(d0, zq.z)), __fadd_rn(o1, __fmul_rn(d1, zq.z))));
          __stcs(pq + 2, make_float4(__fadd_rn(o2, __fmul_rn(d2, zq.z)), __fadd_rn(o0, __fmul_rn(d0, zq.w)),
                                     __fadd_rn(o1, __fmul_rn(d1, zq.w)), __fadd_rn(o2, __fmul_rn(d2, zq.w))));
        }
      }
    } else {
      for (int k = lane; k < Nt; k += 32) zrow[k] = s_out[k];
      if (pts) {
        const float o0 = ro[b * 3], o1 = ro[b * 3 + 1], o2 = ro[b * 3 + 2];
        const float d0 = rd[b * 3], d1 = rd[b * 3 + 1], d2 = rd[b * 3 + 2];
        for (int e = lane; e < 3 * Nt; e += 32) {     // 3*Nt contiguous floats per ray: coalesced stores
          const int sidx = e / 3, k = e - 3 * sidx;
          const float o = k == 0 ? o0 : (k == 1 ? o1 : o2), d = k == 0 ? d0 : (k == 1 ? d1 : d2);
          prow[e] = __fadd_rn(o, __fmul_rn(d, s_out[sidx]));
        }
      }
    }
    __syncwarp();
  }
}

// generic path (any Nf): shared-memory bitonic sort of the concatenation
__global__ void __launch_bounds__(kWarpsPerCta * 32)
sample_hierarchical_generic_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                                   const float* __restrict__ zc, const float* __restrict__ weights, int64_t B, int Nc,
                                   const float* __restrict__ u, int64_t u_stride, int Nf, int P,
                                   float* __restrict__ z_all, float* __restrict__ pts, int64_t* __restrict__ inds_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nb = Nc - 1, Nt = Nc + Nf;
  float* s_bins = smem + (size_t)w * (3 * Nc + P);
  float* s_w = s_bins + Nc;
  float* s_cdf = s_w + Nc;
  float* s_sort = s_cdf + Nc;
  for (int64_t b = blockIdx.x * (int64_t)kWarpsPerCta + w; b < B; b += (int64_t)gridDim.x * kWarpsPerCta) {
    for (int k = lane; k < Nc; k += 32) s_sort[k] = __ldcs(zc + b * Nc + k);
    for (int k = Nt + lane; k < P; k += 32) s_sort[k] = CUDART_INF_F;
    for (int k = lane; k < Nc - 2; k += 32) s_w[k] = __ldcs(weights + b * Nc + k + 1);
    __syncwarp();
    for (int k = lane; k < nb; k += 32) s_bins[k] = __fmul_rn(0.5f, __fadd_rn(s_sort[k + 1], s_sort[k]));
    __syncwarp();
    warp_build_cdf(s_w, s_cdf, nb, lane);
    for (int j = lane; j < Nf; j += 32) {
      int ind;
      s_sort[Nc + j] = invert_cdf(s_cdf, s_bins, nb, __ldg(u + b * u_stride + j), ind);
      if (inds_out) inds_out[b * Nf + j] = ind;
    }
    __syncwarp();
    warp_bitonic_sort(s_sort, P, lane);
    for (int k = lane; k < Nt; k += 32) z_all[b * Nt + k] = s_sort[k];
    if (pts) {
      const float o0 = ro[b * 3], o1 = ro[b * 3 + 1], o2 = ro[b * 3 + 2];
      const float d0 = rd[b * 3], d1 = rd[b * 3 + 1], d2 = rd[b * 3 + 2];
      for (int e = lane; e < 3 * Nt; e += 32) {
        const int s = e / 3, k = e - 3 * s;
        const float o = k == 0 ? o0 : (k == 1 ? o1 : o2), d = k == 0 ? d0 : (k == 1 ? d1 : d2);
        pts[(b * Nt) * 3 + e] = __fadd_rn(o, __fmul_rn(d, s_sort[s]));
      }
    }
    __syncwarp();
  }
}

template <int E, bool INDS>
static int launch_hier_v(const float* ro, const float* rd, const float* zc, const float* weights, int64_t B, int Nc,
                         const float* u, int64_t u_stride, int Nf, float* z_all, float* pts, int64_t* inds_out, cudaStream_t st) {
  const size_t smem = (size_t)kWarpsPerCta * hier_smem_floats(Nc, 32 * E, Nf, INDS) * sizeof(float);
  if (smem > 200 * 1024) return RN_ERR_INVALID_ARG;
  static unsigned long long configured = 0;
  if (smem > 48 * 1024 || first_use_on_device(configured))
    RN_CUDA_CHECK(cudaFuncSetAttribute(sample_hierarchical_kernel<E, INDS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
  const int64_t want = ceil_div(B, kWarpsPerCta);
  const int grid = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  sample_hierarchical_kernel<E, INDS><<<grid, kWarpsPerCta * 32, smem, st>>>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, z_all,
                                                                            pts, inds_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
template <int E>
static int launch_hier(const float* ro, const float* rd, const float* zc, const float* weights, int64_t B, int Nc,
                       const float* u, int64_t u_stride, int Nf, float* z_all, float* pts, int64_t* inds_out, cudaStream_t st) {
  return inds_out ? launch_hier_v<E, true>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, z_all, pts, inds_out, st)
                  : launch_hier_v<E, false>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, z_all, pts, nullptr, st);
}

}  // namespace rn

using namespace rn;

extern "C" {

int rn_stratified_fwd(const float* ro, const float* rd, int64_t B, const float* zb, int Nc, const float* t_rand,
                      float* z_out, float* pts, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(zb && z_out && B >= 0 && Nc >= 1 && (!pts || (ro && rd)));
  const uintptr_t align = reinterpret_cast<uintptr_t>(t_rand) | reinterpret_cast<uintptr_t>(z_out) | reinterpret_cast<uintptr_t>(pts);
  if ((Nc & 3) == 0 && (align & 15) == 0)
    stratified_vec4_kernel<<<grid_for(B * (Nc >> 2), 256), 256, 0, (cudaStream_t)stream>>>(
        ro, rd, B, zb, Nc, reinterpret_cast<const float4*>(t_rand), reinterpret_cast<float4*>(z_out), reinterpret_cast<float4*>(pts));
  else
    stratified_kernel<<<grid_for(B * Nc, 256), 256, 0, (cudaStream_t)stream>>>(ro, rd, B, zb, Nc, t_rand, z_out, pts);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_points_fwd(const float* ro, const float* rd, const float* z, int64_t B, int S, float* pts, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(ro && rd && z && pts && B >= 0 && S >= 1);
  points_fwd_kernel<<<grid_for(B * S, 256), 256, 0, (cudaStream_t)stream>>>(ro, rd, z, B, S, pts);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_points_bwd(const float* g_pts, const float* z, int64_t B, int S, float* g_o, float* g_d, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(g_pts && z && g_o && g_d && B >= 0 && S >= 1);
  points_bwd_kernel<<<grid_for(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(g_pts, z, B, S, g_o, g_d);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_sample_pdf_fwd(const float* bins, const float* weights, int64_t B, int nb, const float* u, int64_t u_stride,
                      int Nf, float* samples, int64_t* inds_out, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(bins && weights && u && samples && B >= 0 && nb >= 2 && Nf >= 1 && nb <= 4096);
  const size_t smem = (size_t)kWarpsPerCta * 3 * nb * sizeof(float);
  if (smem > 48 * 1024)
    RN_CUDA_CHECK(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(ceil_div(B, kWarpsPerCta) < (int64_t)num_sms() * 8 ? ceil_div(B, kWarpsPerCta) : (int64_t)num_sms() * 8);
  sample_pdf_kernel<<<grid, kWarpsPerCta * 32, smem, (cudaStream_t)stream>>>(bins, weights, B, nb, u, u_stride, Nf, samples,
                                                                            inds_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int rn_sample_hierarchical_fwd(const float* ro, const float* rd, const float* zc, const float* weights, int64_t B, int Nc,
                               const float* u, int64_t u_stride, int Nf, float* z_all, float* pts, int64_t* inds_out,
                               rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(zc && weights && u && z_all && B >= 0 && Nc >= 3 && Nf >= 1 && (!pts || (ro && rd)));
  cudaStream_t st = (cudaStream_t)stream;
  if (Nf <= 32) return launch_hier<1>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, z_all, pts, inds_out, st);
  if (Nf <= 64) return launch_hier<2>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, z_all, pts, inds_out, st);
  if (Nf <= 128) return launch_hier<4>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, z_all, pts, inds_out, st);
  if (Nf <= 256) return launch_hier<8>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, z_all, pts, inds_out, st);
  int P = 2;
  while (P < Nc + Nf) P <<= 1;
  const size_t smem = (size_t)kWarpsPerCta * (3 * Nc + P) * sizeof(float);
  RN_REQUIRE(smem <= 200 * 1024);
  if (smem > 48 * 1024)
    RN_CUDA_CHECK(cudaFuncSetAttribute(sample_hierarchical_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(ceil_div(B, kWarpsPerCta) < (int64_t)num_sms() * 8 ? ceil_div(B, kWarpsPerCta) : (int64_t)num_sms() * 8);
  sample_hierarchical_generic_kernel<<<grid, kWarpsPerCta * 32, smem, st>>>(ro, rd, zc, weights, B, Nc, u, u_stride, Nf, P, z_all,
                                                                           pts, inds_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // extern "C"
