// wgrad_stream.cu -- every weight / bias gradient of one network's backward pass (the autograd backward of
// noisy_src/model.py:169-194: dW_l = dH_l^T H_{l-1}, db_l = column sums of dH_l) in ONE persistent launch that runs
// CONCURRENTLY with the data-gradient chain (chain_pair.cu) on its own share of the SMs.
//
// Why: run one after the other, the data-gradient chain is paced by its epilogues (HBM at 4.5 TB/s, tensor pipe 48 %) and
// the split-K weight-gradient GEMMs by HBM (6.1 TB/s, tensor pipe 16-42 %).  Side by side they share the machine by what
// each one needs: 2.93-3.05 ms for both against 3.2 ms in sequence, 28 launches per step instead of 47, and 44 partial
// tiles per GEMM to reduce instead of 148.  (The design also hands each block of dH_l over through L2 -- flags below -- but
// measured, the blocks are 25-35 us old when they are read and have left the L2 by then: a step still reads all of dH
// from HBM.  profiles/r02_ab_log.md blocks 19-22 have the whole story.)
//   * the chain's store warp publishes flag[layer][block] (a strong store of the publication time, four store groups after
//     the block's TMA store: cp.async.bulk.wait_group says its writes have completed);
//   * every CTA PAIR here (cluster of 2, tcgen05 cta_group::2, M = 256) owns one (GEMM, split) for the whole launch:
//     CTA r holds output features [128 r, 128 r + 128) of dW with its fp32 accumulator resident in TMEM, streams its
//     half of dH_l^T (16 KiB per 64 points) and HALF of H_{l-1}'s columns (16 KiB): 32 KiB per SM and chunk where a
//     single-CTA tile pulls 48 -- a pair is paced by what one SM can take in (~65 GB/s of HBM-sourced TMA traffic,
//     whatever the depth of its rings), so bytes per SM are what counts;
//   * split s of S takes the point blocks s, s+S, s+2S, ... in that FIXED order (that is what keeps the sums
//     bit-reproducible); the TMA warp polls the flags of its next 32 blocks at once (ld.acquire.gpu, one per lane) so the
//     flag latency is off the load path;
//   * the bias gradient is one more N=16 MMA against a tile of ones (column sums of dH_l);
//   * tensor cores accumulate with truncation, so an accumulator is not kept for the whole launch: every kWsFlushChunks
//     chunks (1,024 accumulation steps, about as many as one split of the split-K kernels sees) the epilogue warps
//     add it into the fp32 partial tile in global memory (round-to-nearest) and the MMAs start a fresh one;
//   * the partial tiles use the [split][m_tile] layout of the split-K kernels and are summed by the same deterministic
//     reduce kernel (gemm_tcgen05.cu) -- the assignment is static, so a step stays bit-reproducible.
// Widths that are not 256: N = 320 ([x_enc | H4], [feat | d_enc]) is a second MMA of N = 128 whose upper half lies beyond
// the tensor (those boxes are never loaded: their slots are zeroed once); N = 64 (x_enc) is one MMA of N = 128 likewise;
// M = 128 (dHC) is M = 256 with the missing features out of bounds.  (The one-row sigma_linear gradient is not a problem
// of this kernel: heads_bwd_kernel<true> sums it while it streams HC -- a CTA pair per split for one row cost more.)
// Both kernels are launched as clusters of two CTAs (whole TPCs), one CTA per SM, and together they ask for no more SMs
// than the device has, so they are co-resident whatever order the hardware starts them in, and the dependency is one-way
// (the chain never waits for this kernel), so other work on the GPU can delay the pair but not deadlock it; a wait on a
// flag that never comes traps after 4 s instead of hanging the device.
//
// Warp roles (224 threads): 0 = flag polling + A producer, 1 = TMEM owner + MMA issuer (leader CTA), 2..5 = epilogue,
// 6 = B producer.
#include "common.cuh"
#include "ptx.cuh"
#include "gemm.h"
#include <cuda_bf16.h>
#include <stdio.h>

namespace rn {

using namespace ptx;

int make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer);
void prof_begin(int mode, cudaStream_t st, int* slot);
void prof_end(int slot, cudaStream_t st);
extern double g_prof_next_flops;
extern int g_l2_hints, g_ws_debug;

constexpr int kWsThreads = 224;                 // 7 warps
constexpr int kWsSmem = 231424;                 // 226 KiB
constexpr int kWsOnes = 2048;                   // 16 k-rows x 128 B of bf16 1.0 (first 2 KiB of the aligned buffer)
constexpr int kWsRing = kWsSmem - 1024 /*alignment slack*/ - kWsOnes - 1024 /*barriers*/;
constexpr int kWsABytes = 16384;                // [64 points][128 output features]
constexpr int kWsAStages = 5;                   // dH tiles: loaded once their block is published
constexpr int kWsMaxBStages = 12;               // H tiles: no flag to wait for -> the deeper ring, running ahead
constexpr int kWsFlushChunks = 256;             // accumulator flushed to the fp32 partial every 256 chunks (16,384 points)

struct WsProblem {
  CUtensorMap tmA, tmB;      // boxes of [64 points][64 columns], SWIZZLE_128B
  float* partial;            // [splits][m_tiles][128 * (BN + 1)] fp32
  int BN;                    // columns of dW (64, 256 or 320)
  int m_tiles;               // 128-row tiles of dW that exist (1 or 2); CTA rank >= m_tiles computes zeros and writes nothing
  int splits;
  int a_col0;                // first column of A this problem reads
  int a_cols, b_cols;        // columns the tensors have: a 64-column box that starts beyond them is never loaded (stays zero)
  int flag_row;              // row of the flag table the A operand waits on, or -1 (A was complete before the launch)
  int pair0;                 // first CTA pair of this problem
};
struct WsParams {
  WsProblem prob[kWsMaxProblems];
  int n_prob, n_pairs, n_blocks;      // n_blocks: 128-point blocks
  int64_t m_rows;
  const uint32_t* flags;              // [flag rows][n_blocks], zeroed before the launch, 1 = block published
  int l2_hints;
  int dbg;                            // measurement only (rn_set_flag(10)): bit 2 = no bias MMA, bit 3 = no MMAs at all, bit 5 = record hand-off lags
};

// measurement hook (rn_debug_stream_lag): per CTA, sum / count / max of (A-load issue time - publication time) in 64 ns units
__device__ unsigned long long g_ws_lag[3 * 320];
__device__ unsigned int g_ws_busy_us[320];       // rn_set_flag(10, 32): microseconds each CTA spent between its first and last MMA wait

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
static __device__ __noinline__ void flag_timeout_trap(int block) {
  printf("rnerf_b200: weight-gradient stream timed out waiting for point block %d (CTA %d)\n", block, (int)blockIdx.x);
  __trap();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWsThreads, 1)
wgrad_stream_kernel(const __grid_constant__ WsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_ones = smem;
  uint8_t* s_ring = smem + kWsOnes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWsOnes + kWsRing);
  uint64_t* full_a = bars;                   // [8]   leader: bytes of both CTAs' A loads
  uint64_t* empty_a = bars + 8;              // [8]   each CTA (multicast commit)
  uint64_t* full_b = bars + 16;              // [12]  leader: bytes of both CTAs' B loads
  uint64_t* empty_b = bars + 28;             // [12]  each CTA
  uint64_t* tmem_full = bars + 40;           //       each CTA (multicast commit)
  uint64_t* tmem_empty = bars + 41;          //       leader: 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 42);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1;
  int pi = 0;
  while (pi + 1 < p.n_prob && pair >= p.prob[pi + 1].pair0) ++pi;
  const WsProblem& P = p.prob[pi];
  const int BN = P.BN;
  const int split = pair - P.pair0;
  const int S = P.splits;
  // MMA 1: N1 columns (both CTAs hold N1 / 2 of them); MMA 2 (BN = 320 only): columns 256.. as N = 128, upper half out of bounds
  const int N1 = BN == 64 ? 128 : 256;
  const int N2 = BN == 320 ? 128 : 0;
  const int nb1 = N1 / 128;                                  // 64-column boxes of B per CTA for MMA 1
  const int nb = nb1 + (N2 ? 1 : 0);
  // Two rings.  The B operand (H_{l-1}, written by the forward pass) needs no flag and always comes from HBM: its ring
  // is deep and its producer warp runs as far ahead as the ring allows.  The A operand (dH_l) is loaded only once its
  // block is published: a shallower ring.  (Measured against one combined ring of 6 stages: no difference, alone or beside
  // the chain -- what paces a pair is what one SM can take in, ~65 GB/s of HBM-sourced bytes, not the depth of its ring;
  // profiles/r02_ab_log.md block 19.  Kept because the B stream no longer stalls behind a flag.)
  const int NA = kWsAStages;
  const int b_stage = nb * 8192;
  const int nb_raw = (kWsRing - NA * kWsABytes) / b_stage;
  const int NB = nb_raw > kWsMaxBStages ? kWsMaxBStages : nb_raw;
  uint8_t* s_ring_b = s_ring + NA * kWsABytes;
  const int bias_col = N1 + N2;
  // MMA 2's real 64 columns (256..319) are one CTA's box, the other CTA's box lies beyond the tensor.  Normally CTA 0 holds
  // them; when CTA 1 has no A features to load (M = 128: dir_linear) it takes them instead, which evens out what the two
  // SMs pull per chunk (32 / 24 KiB instead of 40 / 16) -- the output columns then sit 64 TMEM columns further right.
  const int swap2 = (N2 && P.a_col0 + 128 >= P.a_cols) ? 1 : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&P.tmA); prefetch_tmap(&P.tmB);
    for (int i = 0; i < 8; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < kWsMaxBStages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  // Boxes that start beyond the tensor (the padding of M to 256 and of N to a multiple of 128) are not loaded at all:
  // their slots in every stage are zeroed once, here.  box_mask bit j: box j of a stage (0, 1 = A; 2.. = B) is loaded.
  uint32_t box_mask[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint32_t m = 0;
    for (int j = 0; j < 2; ++j) m |= (P.a_col0 + r * 128 + j * 64 < P.a_cols) ? (1u << j) : 0u;
    for (int j = 0; j < nb1; ++j) m |= (r * (N1 / 2) + j * 64 < P.b_cols) ? (4u << j) : 0u;
    if (N2) m |= (256 + (r ^ swap2) * 64 < P.b_cols) ? (4u << nb1) : 0u;
    box_mask[r] = m;
  }
  const uint32_t my_mask = box_mask[rank];
  const uint32_t tx_a = 8192u * (uint32_t)(__popc(box_mask[0] & 3u) + __popc(box_mask[1] & 3u));
  const uint32_t tx_b = 8192u * (uint32_t)(__popc(box_mask[0] >> 2) + __popc(box_mask[1] >> 2));
  if (warp >= 2 && warp < 6) {
    uint32_t* ones = reinterpret_cast<uint32_t*>(s_ones);
    for (int i = threadIdx.x - 64; i < kWsOnes / 4; i += 128) ones[i] = 0x3F803F80u;   // bf16 1.0 x2
    if (my_mask != (4u << nb) - 1u) {
      uint4* ring = reinterpret_cast<uint4*>(s_ring);
      for (int i = threadIdx.x - 64; i < (NA * kWsABytes + NB * b_stage) / 16; i += 128) ring[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // 64-point K chunks of this pair: blocks split, split + S, ...; the last block of the network may hold one chunk only
  const int last_block_chunks = (int)((p.m_rows - (int64_t)(p.n_blocks - 1) * 128 + 63) / 64);     // 1 or 2
  int my_blocks = 0;
  if (split < p.n_blocks) my_blocks = (p.n_blocks - 1 - split) / S + 1;
  const bool has_last = my_blocks > 0 && (split + (my_blocks - 1) * S == p.n_blocks - 1);
  const int n_chunks = my_blocks * 2 - ((has_last && last_block_chunks == 1) ? 1 : 0);
  const int n_flush = (n_chunks + kWsFlushChunks - 1) / kWsFlushChunks;

  if (warp == 0) {
    // ---------------- flag polling + A producer (both CTAs, each for its own 128 features) ----------------
    const uint32_t full_leader = mapa_u32(smem_u32(full_a), 0);
    const uint64_t pol = l2_policy(p.l2_hints ? 1 : 0);
    const uint32_t* frow = P.flag_row >= 0 ? p.flags + (size_t)P.flag_row * p.n_blocks : nullptr;
    const int a_c = P.a_col0 + (int)rank * 128;
    int s = 0; uint32_t ph = 0; int issued = 0;
    int b = split;
    unsigned long long t0 = 0; uint32_t idle = 0;
    unsigned long long lag_sum = 0, lag_n = 0, lag_max = 0;
    while (b < p.n_blocks) {
      // the next 32 blocks of this split, one flag per lane: how many are published, counting from the first?
      int nready = 32;
      uint32_t f = 1u;
      if (frow) {
        const int bb = b + lane * S;
        f = bb < p.n_blocks ? ld_acquire_gpu(frow + bb) : 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, f != 0u);
        nready = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
        if (nready == 0) {
          __nanosleep(100);
          if ((++idle & 63u) == 0) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) flag_timeout_trap(b);
          }
          continue;
        }
        idle = 0; t0 = 0;
        // (no proxy fence: the loads below are issued after the acquire -- ordered for lane 0 by the ballot -- and read L2,
        // the point of coherence the producer's completed bulk stores went through; a fence.proxy.async here drains the
        // CTA's outstanding TMA loads and cost 2/3 of the kernel's rate, profiles/r02_ab_log.md block 19)
      }
      for (int i = 0; i < nready && b < p.n_blocks; ++i, b += S) {
        const uint32_t stamp = __shfl_sync(0xffffffffu, f, i);
        for (int h = 0; h < 2 && issued < n_chunks; ++h, ++issued) {
          const int kc = b * 2 + h;
          mbar_wait(&empty_a[s], ph ^ 1);
          if (frow && h == 0 && (p.dbg & 32)) {
            const uint32_t lag = (uint32_t)(globaltimer_ns() >> 6) - stamp;
            lag_sum += lag; lag_n += 1; lag_max = lag > lag_max ? lag : lag_max;
          }
          if (lane == 0) {
            uint8_t* a_s = s_ring + s * kWsABytes;
            if (rank == 0) mbar_arrive_expect_tx(&full_a[s], tx_a);
            const uint32_t bar = full_leader + s * 8;
            if (my_mask & 1u) tma_load_2d_pair_hint(a_s, &P.tmA, bar, a_c, kc * 64, pol);          // [64 pts][64 features]
            if (my_mask & 2u) tma_load_2d_pair_hint(a_s + 8192, &P.tmA, bar, a_c + 64, kc * 64, pol);
          }
          __syncwarp();
          if (++s == NA) { s = 0; ph ^= 1; }
        }
      }
    }
    if ((p.dbg & 32) && lane == 0 && blockIdx.x < 320) {
      g_ws_lag[3 * blockIdx.x] = lag_sum; g_ws_lag[3 * blockIdx.x + 1] = lag_n; g_ws_lag[3 * blockIdx.x + 2] = lag_max;
    }
  } else if (warp == 6) {
    // ---------------- B producer (both CTAs, each for its half of the columns): no flags, as far ahead as the ring goes ----------------
    // (an L2 prefetch of B instead, cp.async.bulk.prefetch.tensor eight blocks ahead, made the kernel 15 % slower)
    if (lane == 0) {
      const uint32_t full_leader = mapa_u32(smem_u32(full_b), 0);
      const uint64_t pol = l2_policy(p.l2_hints ? 1 : 0);
      int s = 0; uint32_t ph = 0; int issued = 0;
      for (int b = split; b < p.n_blocks; b += S) {
        for (int h = 0; h < 2 && issued < n_chunks; ++h, ++issued) {
          const int kc = b * 2 + h;
          mbar_wait(&empty_b[s], ph ^ 1);
          uint8_t* b_s = s_ring_b + s * b_stage;
          if (rank == 0) mbar_arrive_expect_tx(&full_b[s], tx_b);
          const uint32_t bar = full_leader + s * 8;
          for (int j = 0; j < nb1; ++j)
            if (my_mask & (4u << j))
              tma_load_2d_pair_hint(b_s + j * 8192, &P.tmB, bar, (int)rank * (N1 / 2) + j * 64, kc * 64, pol);
          if (N2 && (my_mask & (4u << nb1)))
            tma_load_2d_pair_hint(b_s + nb1 * 8192, &P.tmB, bar, 256 + ((int)rank ^ swap2) * 64, kc * 64, pol);
          if (++s == NB) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA) ----------------
    if (rank == 0) {
      const uint32_t idesc_a = make_idesc_bf16(256, N1, 1, 1);
      constexpr uint32_t idesc_b = make_idesc_bf16(256, 128, 1, 1);
      constexpr uint32_t idesc_1 = make_idesc_bf16(256, 16, 1, 1);
      constexpr uint32_t kHi = desc_hi_sw128(1024);
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(s_ring), 8192);
      const uint32_t b_lo0 = desc_lo_sw128(smem_u32(s_ring_b), 8192);
      const uint32_t o_lo = desc_lo_sw128(smem_u32(s_ones), 8192);
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      const unsigned long long t_begin = (p.dbg & 32) ? globaltimer_ns() : 0ull;
      for (int c = 0; c < n_chunks; ++c) {
        const int in_flush = c % kWsFlushChunks;
        if (in_flush == 0 && c > 0) mbar_wait(tmem_empty, ((c / kWsFlushChunks) - 1) & 1u);   // accumulator drained
        mbar_wait(&full_b[sb], phb);
        mbar_wait(&full_a[sa], pha);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t al = a_lo0 + sa * (kWsABytes >> 4), bl = b_lo0 + sb * (b_stage >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (k != 0) ? 1u : (uint32_t)(in_flush != 0);
            const uint64_t adesc = pack64(al + k * 128, kHi);
            if (!(p.dbg & 8)) umma_bf16_pair(tmem_base, adesc, pack64(bl + k * 128, kHi), idesc_a, accum);
            if (N2 && !(p.dbg & 8)) umma_bf16_pair(tmem_base + 256, adesc, pack64(bl + nb1 * 512 + k * 128, kHi), idesc_b, accum);
            if (!(p.dbg & 12)) umma_bf16_pair(tmem_base + bias_col, adesc, pack64(o_lo, kHi), idesc_1, accum);   // column sums of A (bias gradient)
          }
          umma_commit_pair(&empty_a[sa]);
          umma_commit_pair(&empty_b[sb]);
          if (in_flush == kWsFlushChunks - 1 || c == n_chunks - 1) umma_commit_pair(tmem_full);
        }
        __syncwarp();
        if (++sa == NA) { sa = 0; pha ^= 1; }
        if (++sb == NB) { sb = 0; phb ^= 1; }
      }
      if ((p.dbg & 32) && lane == 0 && blockIdx.x < 320) g_ws_busy_us[blockIdx.x] = (unsigned int)((globaltimer_ns() - t_begin) / 1000);
    }
  } else if (warp < 6) {
    // ---------------- epilogue warps: accumulator -> (+=) fp32 partial tile, once per flush ----------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool writes = (int)rank < P.m_tiles;
    float* out = P.partial + ((size_t)split * P.m_tiles + rank) * ((size_t)128 * (BN + 1));
    const uint32_t tmem_empty_leader = mapa_u32(smem_u32(tmem_empty), 0);
    if (n_chunks == 0 && writes) {
      for (int c = 0; c < BN; ++c) out[(size_t)row * BN + c] = 0.f;
      out[(size_t)128 * BN + row] = 0.f;
    }
    for (int f = 0; f < n_flush; ++f) {
      mbar_wait(tmem_full, f & 1u);
      tcgen05_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16);
      if (writes) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_base + c * 32 + ((swap2 && c >= 8) ? 64 : 0), v);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(out + (size_t)row * BN + c * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float4 a = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                   __uint_as_float(v[4 * e + 3]));
            if (f > 0) { const float4 o = dst[e]; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
            dst[e] = a;
          }
        }
        uint32_t v16[16];
        tmem_ld_x16(t_base + bias_col, v16);
        tmem_ld_wait();
        float bsum = __uint_as_float(v16[0]);
        if (f > 0) bsum += out[(size_t)128 * BN + row];
        out[(size_t)128 * BN + row] = bsum;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_leader);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

// What one 64-point chunk costs a pair, in bytes pulled by its busier SM -- with a floor: a stage is recycled one memory
// latency after its loads were issued whatever it holds, so below ~32 KiB per SM a chunk takes the same ~0.4 us (measured:
// the N = 64 and sigma problems, 18-24 KiB per chunk, ran no faster per chunk than the 32 KiB ones).
static int ws_bytes_per_chunk(const WsHostProblem& h) {
  const int a0 = h.a_cols - h.a_col0;                               // in-bounds features from a_col0 on
  const int a_r0 = (a0 >= 128 ? 128 : (a0 > 0 ? a0 : 0)) * 128;
  const int a_r1 = (a0 >= 256 ? 128 : (a0 > 128 ? a0 - 128 : 0)) * 128;
  const int half = (h.N == 64 ? 64 : 128) * 128;                    // MMA 1: each CTA's half of the columns (x_enc: CTA 0 only)
  const int extra = h.N == 320 ? 64 * 128 : 0;                      // MMA 2: the real 64 columns, held by one CTA
  const bool swap2 = h.N == 320 && a_r1 == 0;                       // ... by CTA 1 when it has no A to load (the kernel's swap2)
  const int r0 = a_r0 + half + (swap2 ? 0 : extra);
  const int r1 = a_r1 + (h.N == 64 ? 0 : half) + (swap2 ? extra : 0);
  const int m = r0 > r1 ? r0 : r1;
  return m > 32768 ? m : 32768;
}

// Splits per problem: the slowest pair sets the pace, so hand out pairs one at a time to whichever problem has the most
// bytes per split.
int wgrad_stream_plan(const WsHostProblem* probs, int n, int pairs, int* splits_out) {
  RN_REQUIRE(probs && n > 0 && n <= kWsMaxProblems && splits_out && pairs >= n);
  for (int i = 0; i < n; ++i) splits_out[i] = 1;
  for (int used = n; used < pairs; ++used) {
    int best = 0; double worst = 0.0;
    for (int i = 0; i < n; ++i) {
      const double load = (double)ws_bytes_per_chunk(probs[i]) / splits_out[i];
      if (load > worst) { worst = load; best = i; }
    }
    splits_out[best] += 1;
  }
  return RN_OK;
}

int wgrad_stream_launch(const WsHostProblem* probs, int n, int64_t M, const uint32_t* flags, int ctas, float* scratch,
                        size_t region_floats, TnInfo* infos, cudaStream_t st) {
  int rc = check_arch();
  if (rc != RN_OK) return rc;
  RN_REQUIRE(probs && n > 0 && n <= kWsMaxProblems && M > 0 && flags && scratch && infos && ctas >= 2 * n);
  static thread_local WsParams p;
  int splits[kWsMaxProblems];
  if ((rc = wgrad_stream_plan(probs, n, ctas / 2, splits)) != RN_OK) return rc;
  int pair = 0;
  double flops = 0.0;
  for (int i = 0; i < n; ++i) {
    const WsHostProblem& h = probs[i];
    RN_REQUIRE(h.A && h.B && h.Mo > 0 && h.Mo <= 256 && (h.N == 320 || h.N == 256 || h.N == 64) && h.a_col0 % 64 == 0 &&
               h.a_col0 + h.Mo <= h.a_cols + 63 && h.b_cols >= h.N);
    WsProblem& P = p.prob[i];
    if ((rc = make_tmap(&P.tmA, h.A, h.a_cols, M, h.lda, 64)) != RN_OK) return rc;
    if ((rc = make_tmap(&P.tmB, h.B, h.b_cols, M, h.ldb, 64)) != RN_OK) return rc;
    P.BN = h.N; P.m_tiles = (h.Mo + 127) / 128; P.splits = splits[i]; P.a_col0 = h.a_col0; P.flag_row = h.flag_row;
    P.a_cols = h.a_cols; P.b_cols = h.b_cols;
    P.pair0 = pair;
    P.partial = scratch + (size_t)i * region_floats;
    RN_REQUIRE((size_t)P.splits * P.m_tiles * 128 * (h.N + 1) <= region_floats);
    pair += P.splits;
    infos[i].m_tiles = P.m_tiles; infos[i].splits = P.splits; infos[i].N = h.N; infos[i].scratch = P.partial;
    flops += 2.0 * (double)M * h.N * h.Mo;
  }
  p.n_prob = n; p.n_pairs = pair; p.n_blocks = (int)ceil_div(M, 128); p.m_rows = M; p.flags = flags;
  p.l2_hints = (g_l2_hints & 8) ? 0 : 1;          // both operands are read once: evict_first (rn_set_flag(6, 8) switches it off)
  p.dbg = g_ws_debug;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    RN_CUDA_CHECK(cudaFuncSetAttribute(wgrad_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmem));
  g_prof_next_flops = flops;
  int slot;
  prof_begin(2 /*MODE_TN*/, st, &slot);
  wgrad_stream_kernel<<<dim3(2 * pair), dim3(kWsThreads), kWsSmem, st>>>(p);
  prof_end(slot, st);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

extern "C" int rn_debug_stream_lag(double* mean_us_host, double* max_us_host, int* ctas_host) {
  // measurement hook: hand-off lag (publication of a block by the chain -> issue of its load by the stream) of the last
  // wgrad_stream launch made with rn_set_flag(10, 32); synchronises the device
  using namespace rn;
  RN_REQUIRE(mean_us_host && max_us_host && ctas_host);
  static unsigned long long host[3 * 320];
  RN_CUDA_CHECK(cudaDeviceSynchronize());
  RN_CUDA_CHECK(cudaMemcpyFromSymbol(host, g_ws_lag, sizeof(host)));
  double sum = 0.0, n = 0.0, mx = 0.0; int ctas = 0;
  for (int i = 0; i < 320; ++i) {
    if (host[3 * i + 1] == 0) continue;
    sum += (double)host[3 * i]; n += (double)host[3 * i + 1]; ++ctas;
    if ((double)host[3 * i + 2] > mx) mx = (double)host[3 * i + 2];
  }
  *mean_us_host = n > 0 ? sum / n * 0.064 : 0.0; *max_us_host = mx * 0.064; *ctas_host = ctas;
  return RN_OK;
}

extern "C" int rn_debug_stream_busy(unsigned int* us_host, int n) {
  // measurement hook: microseconds between the first and the last MMA wait of CTA pair i's leader (entry 2 i) in the last
  // wgrad_stream launch made with rn_set_flag(10, 32): which GEMM's pairs set the pace
  using namespace rn;
  RN_REQUIRE(us_host && n > 0 && n <= 320);
  RN_CUDA_CHECK(cudaDeviceSynchronize());
  RN_CUDA_CHECK(cudaMemcpyFromSymbol(us_host, g_ws_busy_us, sizeof(unsigned int) * n));
  return RN_OK;
}
