// umma_bench.cu -- microbenchmark: what does the tensor pipe sustain with both operands in shared memory?
// Compares tcgen05.mma cta_group::1 (M=128 per CTA) with cta_group::2 (M=256 per CTA pair, B split across the pair),
// N = 256 / 128, K = 16 per instruction, SWIZZLE_128B K-major operands, and checks the pair layout numerically.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench scripts/umma_bench.cu && ./umma_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (clock64() - t0 > 2000000000LL) { printf("timeout\n"); __trap(); }
  }
}

constexpr int kStages = 4;

// D[128 x BN] per CTA.  A: [128 rows][64 k] SW128 K-major (16 KB / stage).  B: [BN/CG rows][64 k].
template <int CG, int BN>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, int commit_every_stage, long long* out_cycles, float* out_d) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kA = 128 * 128, kB = (BN / CG) * 128;
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + kStages * kA;
  __shared__ uint64_t bar_done, bar_stage;
  __shared__ uint32_t tmem_slot;
  const uint32_t rank = (CG == 2) ? cluster_rank() : 0;
  const int warp = threadIdx.x >> 5;

  // fill operands: A[r][k] = ((3 gr + 5 k + s) % 7) - 3, B[n][k] = ((2 gn + 3 k + s) % 5) - 2   (exact in bf16 / fp32)
  for (int s = 0; s < kStages; ++s) {
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
      const int r = i / 64, k = i % 64, gr = rank * 128 + r;
      const float v = (float)((3 * gr + 5 * k + s) % 7 - 3);
      *reinterpret_cast<__nv_bfloat16*>(s_a + s * kA + r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(v);
    }
    for (int i = threadIdx.x; i < (BN / CG) * 64; i += blockDim.x) {
      const int n = i / 64, k = i % 64, gn = rank * (BN / CG) + n;
      const float v = (float)((2 * gn + 3 * k + s) % 5 - 2);
      *reinterpret_cast<__nv_bfloat16*>(s_b + s * kB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(v);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_stage)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  long long cycles = 0;
  if (threadIdx.x == 0 && rank == 0) {
    constexpr uint32_t idesc = idesc_bf16(128 * CG, BN);
    const long long t0 = clock64();
    uint32_t stage_commits = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (it & 1) * 256;
#pragma unroll
      for (int s = 0; s < kStages; ++s) {
        const uint32_t a_addr = smem_u32(s_a + s * kA), b_addr = smem_u32(s_b + s * kB);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma<CG>(d, desc_sw128(a_addr + k * 32, 16, 1024), desc_sw128(b_addr + k * 32, 16, 1024), idesc, (s | k) != 0);
        if (commit_every_stage) { commit<CG>(&bar_stage); ++stage_commits; }
      }
    }
    commit<CG>(&bar_done);
    mbar_wait(&bar_done, 0);
    cycles = clock64() - t0;
    (void)stage_commits;
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar_done, 0);      // multicast commit arrives here too
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0 && rank == 0) out_cycles[blockIdx.x] = cycles;

  // dump the accumulator of the last iteration (only block/cluster 0)
  if (out_d && blockIdx.x < CG) {
    const uint32_t d = tmem_base + ((iters - 1) & 1) * 256 + ((uint32_t)(warp * 32) << 16);
    const int row = rank * 128 + warp * 32 + (threadIdx.x & 31);
    for (int c = 0; c < BN; c += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(d + c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int e = 0; e < 16; ++e) out_d[(size_t)row * BN + c + e] = __uint_as_float(v[e]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <int CG, int BN>
static void run(int iters, int commit_every_stage, int grid) {
  constexpr int smem = kStages * (128 * 128 + (BN / CG) * 128) + 1024;
  auto kern = bench_kernel<CG, BN>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d_cycles; float* d_out;
  CK(cudaMalloc(&d_cycles, grid * sizeof(long long)));
  CK(cudaMemset(d_cycles, 0, grid * sizeof(long long)));
  CK(cudaMalloc(&d_out, (size_t)128 * CG * BN * sizeof(float)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  // correctness: one iteration
  CK(cudaLaunchKernelEx(&cfg, kern, 1, 0, d_cycles, d_out));
  CK(cudaDeviceSynchronize());
  std::vector<float> h((size_t)128 * CG * BN);
  CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  double max_err = 0;
  for (int r = 0; r < 128 * CG; ++r)
    for (int n = 0; n < BN; ++n) {
      double ref = 0;
      for (int s = 0; s < kStages; ++s)
        for (int k = 0; k < 64; ++k) ref += (double)((3 * r + 5 * k + s) % 7 - 3) * (double)((2 * n + 3 * k + s) % 5 - 2);
      double e = fabs(ref - h[(size_t)r * BN + n]);
      if (e > max_err) max_err = e;
    }
  // timing
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, kern, iters, commit_every_stage, d_cycles, (float*)nullptr));
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, kern, iters, commit_every_stage, d_cycles, (float*)nullptr));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  const double n_mma = (double)iters * kStages * 4;
  const double flops = 2.0 * 128 * BN * 16 * n_mma * grid;          // per CTA: 128 rows x BN x 16 per instruction
  printf("cta_group::%d N=%d commit/stage=%d grid=%d: max|err|=%g  cycles/MMA=%.1f  wall %.3f ms  %.1f TFLOP/s  (SM clock ~%.0f MHz)\n",
         CG, BN, commit_every_stage, grid, max_err, (double)cmax / n_mma, ms, flops / ms * 1e-9, (double)cmax / ms * 1e-3);
  CK(cudaFree(d_cycles)); CK(cudaFree(d_out));
}


// ---- handshake variant (cta_group::1): the real pipeline's barrier traffic without any data movement ----
// warp 1 lane 0 = "producer": wait empty[s] -> arrive full[s];  warp 0 lane 0 = MMA issuer: wait full[s] -> 4 MMAs -> commit empty[s].
// every `tile_kc` stages the accumulator flips and (optionally) a tmem_full commit + tmem_empty wait with an "epilogue" thread.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int NS>
__global__ void __launch_bounds__(128, 1) handshake_kernel(int n_kc, int tile_kc, int n_commits, int use_epi, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kA = 128 * 128, kB = 256 * 128;
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + 2 * kA;
  __shared__ uint64_t full[8], empty[8], empty2[8], tfull[2], tempty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * kA + 2 * kB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 0xff);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty2[i])));
    }
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tfull[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tempty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = idesc_bf16(128, 256);
    const long long t0 = clock64();
    int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0;
    for (int kc = 0; kc < n_kc; ++kc) {
      const int kin = kc % tile_kc;
      if (kin == 0 && use_epi) { mbar_wait(&tempty[acc], acc_ph ^ 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(s_a + (s & 1) * kA), b_addr = smem_u32(s_b + (s & 1) * kB);
      const uint32_t d = tmem_base + acc * 256;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma<1>(d, desc_sw128(a_addr + k * 32, 16, 1024), desc_sw128(b_addr + k * 32, 16, 1024), idesc, (kin | k) != 0);
      commit<1>(&empty[s]);
      if (n_commits >= 2) commit<1>(&empty2[s]);
      if (kin == tile_kc - 1) {
        if (use_epi) commit<1>(&tfull[acc]);
        acc ^= 1; if (acc == 0) acc_ph ^= 1;
      }
      if (++s == NS) { s = 0; ph ^= 1; }
    }
    // drain: wait for the last stage's release
    const int last = (s + NS - 1) % NS;
    mbar_wait(&empty[last], (s == 0) ? (ph ^ 1) : ph);
    out_cycles[blockIdx.x] = clock64() - t0;
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int kc = 0; kc < n_kc; ++kc) {
      mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive(&full[s]);
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp == 2 && lane == 0 && use_epi) {
    int acc = 0; uint32_t acc_ph = 0;
    for (int t = 0; t < n_kc / tile_kc; ++t) {
      mbar_wait(&tfull[acc], acc_ph);
      mbar_arrive(&tempty[acc]);
      acc ^= 1; if (acc == 0) acc_ph ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int NS>
static void run_hs(int n_kc, int tile_kc, int n_commits, int use_epi, int grid) {
  constexpr int smem = 2 * 128 * 128 + 2 * 256 * 128 + 1024;
  auto kern = handshake_kernel<NS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, grid * sizeof(long long)));
  kern<<<grid, 128, smem>>>(n_kc, tile_kc, n_commits, use_epi, d_cycles);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  kern<<<grid, 128, smem>>>(n_kc, tile_kc, n_commits, use_epi, d_cycles);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  printf("handshake ring=%d tile_kc=%d commits/kc=%d epi=%d: cycles/MMA=%.1f  wall %.3f ms  %.1f TFLOP/s\n", NS, tile_kc, n_commits,
         use_epi, (double)cmax / (n_kc * 4.0), ms, 2.0 * 128 * 256 * 16 * 4.0 * n_kc * grid / ms * 1e-9);
  CK(cudaFree(d_cycles));
}

// ---- lean issue loop: whole warp converged, elect.sync around the issue, descriptors as (lo + const, hi const) ----
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@!P bra WAIT_LOOP;\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
template <int NS>
__global__ void __launch_bounds__(128, 1) lean_kernel(int n_kc, int tile_kc, int n_commits, int use_epi, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kA = 128 * 128, kB = 256 * 128;
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + 2 * kA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kA + 2 * kB);
  uint64_t* full = bars; uint64_t* empty = bars + 8; uint64_t* empty2 = bars + 16; uint64_t* tfull = bars + 24; uint64_t* tempty = bars + 26;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * kA + 2 * kB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 0xff);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 28; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = idesc_bf16(128, 256);
    constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(s_a) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(s_b) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), empty20 = smem_u32(empty2), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
    const long long t0 = clock64();
    int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0; int kin = 0;
    for (int kc = 0; kc < n_kc; ++kc) {
      if (kin == 0 && use_epi) { mbar_wait_fast(tempty0 + acc * 8, acc_ph ^ 1); }
      mbar_wait_fast(full0 + s * 8, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t al = a_lo0 + (s & 1) * (kA >> 4), bl = b_lo0 + (s & 1) * (kB >> 4);
        const uint32_t d = tmem_base + acc * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) mma<1>(d, pack64(al + 2 * k, hi), pack64(bl + 2 * k, hi), idesc, (kin | k) != 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + s * 8) : "memory");
        if (n_commits >= 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty20 + s * 8) : "memory");
        if (kin == tile_kc - 1 && use_epi) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tfull0 + acc * 8) : "memory");
      }
      __syncwarp();
      if (++kin == tile_kc) { kin = 0; acc ^= 1; if (acc == 0) acc_ph ^= 1; }
      if (++s == NS) { s = 0; ph ^= 1; }
    }
    const int last = (s + NS - 1) % NS;
    mbar_wait_fast(empty0 + last * 8, (s == 0) ? (ph ^ 1) : ph);
    if (lane == 0) out_cycles[blockIdx.x] = clock64() - t0;
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int kc = 0; kc < n_kc; ++kc) {
      mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive(&full[s]);
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp == 2 && lane == 0 && use_epi) {
    int acc = 0; uint32_t acc_ph = 0;
    for (int t = 0; t < n_kc / tile_kc; ++t) {
      mbar_wait(&tfull[acc], acc_ph);
      mbar_arrive(&tempty[acc]);
      acc ^= 1; if (acc == 0) acc_ph ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}
template <int NS>
static void run_lean(int n_kc, int tile_kc, int n_commits, int use_epi, int grid) {
  constexpr int smem = 2 * 128 * 128 + 2 * 256 * 128 + 1024 + 512;
  auto kern = lean_kernel<NS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, grid * sizeof(long long)));
  kern<<<grid, 128, smem>>>(n_kc, tile_kc, n_commits, use_epi, d_cycles);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  kern<<<grid, 128, smem>>>(n_kc, tile_kc, n_commits, use_epi, d_cycles);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  printf("lean      ring=%d tile_kc=%d commits/kc=%d epi=%d: cycles/MMA=%.1f  wall %.3f ms  %.1f TFLOP/s\n", NS, tile_kc, n_commits,
         use_epi, (double)cmax / (n_kc * 4.0), ms, 2.0 * 128 * 256 * 16 * 4.0 * n_kc * grid / ms * 1e-9);
  CK(cudaFree(d_cycles));
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int iters = 20000;
  run<1, 256>(iters, 0, sms);
  run<1, 256>(iters, 1, sms);
  run<2, 256>(iters, 0, sms);
  run<2, 256>(iters, 1, sms);
  run<1, 128>(iters, 0, sms);
  run<2, 128>(iters, 0, sms);
  run<1, 256>(iters, 0, 1);
  run<2, 256>(iters, 0, 2);
  const int n_kc = 40000;
  run_hs<1>(n_kc, 4, 1, 0, sms);
  run_hs<2>(n_kc, 4, 1, 0, sms);
  run_hs<3>(n_kc, 4, 1, 0, sms);
  run_hs<4>(n_kc, 4, 1, 0, sms);
  run_hs<8>(n_kc, 4, 1, 0, sms);
  run_hs<2>(n_kc, 4, 2, 0, sms);
  run_hs<2>(n_kc, 4, 2, 1, sms);
  run_hs<4>(n_kc, 4, 2, 1, sms);
  run_hs<8>(n_kc, 4, 2, 1, sms);
  run_hs<2>(n_kc, 1, 2, 1, sms);
  run_lean<2>(n_kc, 4, 1, 0, sms);
  run_lean<4>(n_kc, 4, 1, 0, sms);
  run_lean<2>(n_kc, 4, 2, 1, sms);
  run_lean<4>(n_kc, 4, 2, 1, sms);
  run_lean<4>(n_kc, 1, 2, 1, sms);
  return 0;
}
