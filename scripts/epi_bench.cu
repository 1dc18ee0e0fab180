// epi_bench.cu -- what does the chain epilogue cost by itself?  16 warps read a 128 x 256 fp32 accumulator from TMEM,
// add a bias, ReLU, convert to bf16 and store the swizzled tile to shared memory; no MMA, no TMA running.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_bench scripts/epi_bench.cu && ./epi_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__constant__ float2 c_bias2[128];
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint64_t fadd2(uint32_t a_lo, uint32_t a_hi, float2 b) {
  uint64_t a, bb, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a_lo), "r"(a_hi));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(bb));
  return r;
}
// VARIANT bits: 1 = scalar FADD instead of FADD2, 2 = truncating PRMT instead of F2FP, 4 = no HMNMX2, 8 = no STS,
// 16 = no fence.proxy.async, 32 = only one LDTM per item consumed (math skipped)
template <int VARIANT>
__global__ void __launch_bounds__(640, 1) epi_kernel(int items, long long* out_cycles, uint32_t* sink, int spin_mode, int mma_depth) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t mma_bar[16];
  __shared__ volatile int epi_done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar[i])));
    epi_done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  uint32_t acc_sink = 0;
  __shared__ float2 s_bias2[128];
  if (threadIdx.x < 128) s_bias2[threadIdx.x] = c_bias2[threadIdx.x];
  __syncthreads();
  if (warp < 16) {
    const int q = warp & 3, cq = warp >> 2, row = q * 32 + lane;
    const long long t0 = clock64();
    for (int it = 0; it < items; ++it) {
      const int slot = mma_depth ? 0 : (it & 1);
      uint8_t* s_tile = smem + slot * 65536;
      const uint32_t t_addr = tmem_base + slot * 256 + ((uint32_t)(q * 32) << 16) + cq * 64;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int G = cq * 2 + g;
        uint32_t v[32];
        tmem_ld_x32(t_addr + g * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (VARIANT & 32) { acc_sink ^= v[0]; continue; }
        uint8_t* box = s_tile + (G >> 1) * 16384 + row * 128;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int lchunk = (G & 1) * 4 + cc;
          uint4* dst = reinterpret_cast<uint4*>(box + ((lchunk ^ (row & 7)) << 4));
          uint32_t packed[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = cc * 4 + e;
            float lo, hi;
            const float2 b = (VARIANT & 64) ? make_float2(0.01f, 0.02f) : (VARIANT & 128) ? c_bias2[(g * 16 + i) & 127] : (VARIANT & 256) ? s_bias2[(G * 16 + i) & 127] : c_bias2[(G * 16 + i) & 127];
            if (VARIANT & 1) { lo = __uint_as_float(v[2 * i]) + b.x; hi = __uint_as_float(v[2 * i + 1]) + b.y; }
            else {
              const uint64_t r = fadd2(v[2 * i], v[2 * i + 1], b);
              asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
            }
            uint32_t pk;
            if (VARIANT & 2) pk = __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
            else { __nv_bfloat162 p2 = __floats2bfloat162_rn(lo, hi); pk = *reinterpret_cast<uint32_t*>(&p2); }
            if (!(VARIANT & 4)) { __nv_bfloat162 p2 = *reinterpret_cast<__nv_bfloat162*>(&pk); p2 = __hmax2(p2, __float2bfloat162_rn(0.f)); pk = *reinterpret_cast<uint32_t*>(&p2); }
            packed[e] = pk;
          }
          if (VARIANT & 8) acc_sink ^= packed[0] ^ packed[1] ^ packed[2] ^ packed[3];
          else *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (!(VARIANT & 16)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
    }
    if (threadIdx.x == 0) { out_cycles[blockIdx.x] = clock64() - t0; epi_done = 1; }
  }
  // concurrent MMA stream into the OTHER accumulator columns (as the other tile slot of the chain kernel), at most
  // mma_depth chunks of 4 MMAs (512 cycles each) in flight
  if (warp == 16 && mma_depth > 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 65536);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int c = 0;
    while (!epi_done && c < 4000000) {
      if (c >= mma_depth) {       // wait for chunk c - mma_depth
        const int j = c - mma_depth; const uint32_t bar = smem_u32(&mma_bar[j & 15]), par = (j >> 4) & 1; uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
      }
      if (elect_one()) {
        const uint32_t d = tmem_base + 256;       // columns 256..511: never read by the epilogue warps
        for (int k = 0; k < 4; ++k)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(d), "l"(desc_sw128(a_addr + k * 32)), "l"(desc_sw128(b_addr + k * 32)), "r"(idesc), "r"(1) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar[c & 15])) : "memory");
      }
      __syncwarp();
      ++c;
    }
  }
  __shared__ uint64_t spin_bar;
  __shared__ volatile int done_flag;
  if (warp >= 17 && spin_mode) {
    // spinning waiters, as the idle MMA / producer warps of the chain kernel: spin_mode 1 = plain try_wait loop (all lanes),
    // 2 = try_wait with a suspend-time hint, 3 = lane 0 only, 4 = try_wait + nanosleep back-off
    if (threadIdx.x == 17 * 32) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&spin_bar))); done_flag = 0; }
    __syncwarp();
    if (spin_mode != 3 || lane == 0) {
      uint32_t ok = 0; int n = 0;
      while (!ok && n < 200000000) {
        if (spin_mode == 2)
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&spin_bar)), "r"(0), "r"(1000000) : "memory");
        else
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&spin_bar)), "r"(0) : "memory");
        if (spin_mode == 4 && !ok) __nanosleep(64);
        ++n;
        if (done_flag) break;
      }
      acc_sink ^= n;
    }
  }
  if (warp == 0 && lane == 0) done_flag = 1;
  if (acc_sink == 0x12345u) sink[0] = acc_sink;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}
template <int VARIANT>
static void run(const char* name, int spin_mode = 0, int mma_depth = 0) {
  const int items = 4000, grid = 148, smem = 2 * 65536 + 1024;
  auto kern = epi_kernel<VARIANT>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d_c; uint32_t* d_s;
  CK(cudaMalloc(&d_c, grid * 8)); CK(cudaMalloc(&d_s, 4));
  kern<<<grid, 640, smem>>>(items, d_c, d_s, spin_mode, mma_depth);
  kern<<<grid, 640, smem>>>(items, d_c, d_s, spin_mode, mma_depth);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), d_c, grid * 8, cudaMemcpyDeviceToHost));
  long long m = 0; for (auto x : c) if (x > m) m = x;
  printf("%-60s %8.1f cycles per 128x256 item\n", name, (double)m / items);
  CK(cudaFree(d_c)); CK(cudaFree(d_s));
}
int main() {
  std::vector<float> b(256, 0.01f);
  CK(cudaMemcpyToSymbol(c_bias2, b.data(), 256 * 4));
  run<0>("full epilogue (FADD2, F2FP, HMNMX2, STS.128, fence)");
  run<1>("scalar FADD");
  run<2>("PRMT truncation instead of F2FP");
  run<4>("no HMNMX2");
  run<8>("no STS");
  run<16>("no fence.proxy.async");
  run<32>("LDTM + wait only");
  run<2 | 4 | 8 | 16>("FADD2 only");
  run<64>("full, bias = immediate (no constant loads)");
  run<128>("full, bias index compile-time (uniform constant operand)");
  run<256>("full, bias from shared memory (LDS.64 broadcast)");
  run<64 | 8>("no constant loads, no STS");
  run<0>("full + 3 warps spinning on try_wait (all lanes)", 1);
  run<0>("full + 3 warps spinning on try_wait with suspend hint", 2);
  run<0>("full + 3 warps spinning on try_wait (lane 0 only)", 3);
  run<0>("full + 3 warps spinning on try_wait + nanosleep(64)", 4);
  run<32>("LDTM + wait only, MMAs in flight: 1 chunk (4 MMAs)", 0, 1);
  run<32>("LDTM + wait only, MMAs in flight: 2 chunks", 0, 2);
  run<32>("LDTM + wait only, MMAs in flight: 4 chunks", 0, 4);
  run<32>("LDTM + wait only, MMAs in flight: 8 chunks", 0, 8);
  run<0>("full epilogue, MMAs in flight: 1 chunk", 0, 1);
  run<0>("full epilogue, MMAs in flight: 2 chunks", 0, 2);
  run<0>("full epilogue, MMAs in flight: 4 chunks", 0, 4);
  run<0>("full epilogue, MMAs in flight: 8 chunks", 0, 8);
  return 0;
}
