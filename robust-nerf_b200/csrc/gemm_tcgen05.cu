// gemm_tcgen05.cu -- the dense contraction of the NeRF MLP on 5th-gen tensor cores.
//
// Replaces the 12 addmm (+ReLU, +cat) of noisy_src/model.py:169-194 and the 24 mm of their
// autograd backward.  One warp-specialised kernel, three operand-layout modes:
//   NT  D[M,N]  = act(A[M,K] * B[N,K]^T + bias)            forward layer        (A, B K-major)
//   NN  D[M,N]  = (A[M,K] * B[K,N]) .* (mask > 0)          data gradient        (B MN-major)
//   TN  D[Mo,N] = A[K,Mo]^T * B[K,N]  (+ column sums of A) weight/bias gradient (A, B MN-major)
// Operands are bf16, staged global->shared by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through an
// mbarrier ring; the MMA is tcgen05.mma (cta_group::1, M=128, N<=256, K=16) issued by one elected
// thread with fp32 accumulators in TMEM; the epilogue reads TMEM with tcgen05.ld, applies
// bias/ReLU (or the ReLU mask), converts to bf16, stages the tile in swizzled shared memory and
// writes it back with a TMA store.  NT/NN are persistent over 128-row tiles with a
// double-buffered TMEM accumulator so tile i's epilogue overlaps tile i+1's MMAs; TN is split-K
// over the point dimension with fp32 partial tiles reduced by a second (deterministic) kernel,
// and gets the bias gradient for free from one extra N=16 MMA against a tile of ones.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include "common.cuh"
#include "ptx.cuh"
#include "gemm.h"
#include "mlp_layout.h"
#include <cuda_bf16.h>
#include <stdio.h>
#include <utility>
#include <vector>

namespace rn {

using namespace ptx;

enum { MODE_NT = 0, MODE_NN = 1, MODE_TN = 2 };

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;               // one 128-byte swizzle span of bf16
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kGemmThreads = 224;         // 7 warps: A producer, MMA, 4 epilogue, B producer
constexpr int kEpiThreads = 128;
constexpr int kOnesBytes = 2048;          // 16 k-rows x 128 B of bf16 1.0 (TN bias-gradient trick)
constexpr int kMaxSmem = 231424;          // 226 KiB (1 KiB under the 227 KiB per-block limit)

struct GemmArgs {
  int m_tiles;            // NT/NN: 128-row tiles of D;  TN: 128-row tiles of Mo
  int k_chunks;           // ceil(K / 64)
  int k_total;            // K
  const float* bias;      // NT (may be null)
  int relu;               // NT
  int64_t m_rows;         // NT/NN: valid rows of D (per-row global accesses are bounds-checked against it)
  uint32_t* mask_out;     // NT: packed ReLU mask of D, [M][BN/32] words (bit j of word c <=> D[m, 32c+j] > 0), or null
  const uint32_t* mask_bits;  // NN: packed mask applied to D (same layout), or null
  int n_heads;            // NT: 0, or number of fused fp32 head dot-products (1 = sigma, 3 = rgb)
  const float* head_w;    // NT: [n_heads][BN] fp32
  const float* head_b;    // NT: [n_heads] fp32
  float* head_out;        // NT: raw[M][4]; head h is written to column head_col + h
  int head_col;
  float* partial;         // TN: [m_tiles*splits][128*(BN+1)] fp32
  int splits;             // TN
  int chunks_per_split;   // TN
  int l2_hints;           // TN: both operands are streamed once -> evict_first (rn_set_flag(6))
};

template <int BN, int MODE>
struct GemmCfg {
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kHeadBytes = (MODE == MODE_NT) ? 3 * BN * 4 : 0;    // fused head weights (fp32)
  static constexpr int kMisc = 2048 /*bias + barriers*/ + 1024 /*alignment slack*/;
  // NT / NN: the activation operand A streams from HBM (needs ~100 KB in flight per SM to cover
  // the latency), the weight operand B comes from L2: separate rings, deep for A, shallow for B.
  // output staging: two buffers of up to 128 columns each, so the TMA store of one half drains while the
  // epilogue converts the next half (or the next tile)
  static constexpr int kHalfCols = BN >= 128 ? 128 : BN;
  static constexpr int kHalfBytes = (kHalfCols / 64) * 16384;
  static constexpr int kStagingBytes = (MODE == MODE_TN) ? kOnesBytes : 2 * kHalfBytes;
  static constexpr int kNB = 2;
  static constexpr int kNARaw = (kMaxSmem - kMisc - kHeadBytes - kStagingBytes - kNB * kBBytes) / kABytes;
  static constexpr int kNA = kNARaw > 8 ? 8 : kNARaw;
  // TN: A and B both stream from HBM, one combined ring
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kNSRaw = (kMaxSmem - kMisc - kStagingBytes) / kStageBytes;
  static constexpr int kNS = kNSRaw > 8 ? 8 : kNSRaw;
  static constexpr int kRingBytes = (MODE == MODE_TN) ? kNS * kStageBytes : kNA * kABytes + kNB * kBBytes;
  static constexpr int kAccCols = (MODE == MODE_TN) ? (BN + 16) : 2 * BN;
  static constexpr int kTmemCols = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : kAccCols <= 128 ? 128 : kAccCols <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kRingBytes + kStagingBytes + kHeadBytes + kMisc;
  static_assert(kNA >= 3 && kNS >= 2, "pipeline too shallow");
  static_assert(kSmemBytes <= kMaxSmem, "shared memory budget exceeded");
  static_assert(BN % 64 == 0 && (BN <= 256 || (MODE == MODE_TN && BN == 320)), "BN: 64..256, or 320 for the TN mode");
};

// HEADS (NT only): number of fp32 head dot-products fused into the epilogue (0, 1 = sigma, 3 = rgb).
// WMASK (NT only): also emit the packed ReLU mask of the output tile (training forward).
template <int BN, int MODE, int HEADS, bool WMASK>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, GemmArgs args) {
  using Cfg = GemmCfg<BN, MODE>;
  constexpr int NA = Cfg::kNA, NB = Cfg::kNB, NS = Cfg::kNS;
  constexpr int NW = BN / 32;                // mask words per row
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // keep the pointer in the shared address space (LDS/STS, not generic LD/ST): offset arithmetic only
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;                                   // NT/NN: A ring;  TN: combined ring
  uint8_t* s_b = smem + NA * kABytes;                    // NT/NN: B ring
  uint8_t* s_staging = smem + Cfg::kRingBytes;
  float* s_head = reinterpret_cast<float*>(s_staging + Cfg::kStagingBytes);           // [3][BN] (NT only)
  float* s_bias = reinterpret_cast<float*>(s_staging + Cfg::kStagingBytes + Cfg::kHeadBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* full_a = bars;                   // [8]
  uint64_t* empty_a = bars + 8;              // [8]
  uint64_t* full_b = bars + 16;              // [2]
  uint64_t* empty_b = bars + 18;             // [2]
  uint64_t* tmem_full = bars + 20;           // [2]
  uint64_t* tmem_empty = bars + 22;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---------------- one-time setup ----------------
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    if (MODE != MODE_TN) prefetch_tmap(&tmD);
    for (int i = 0; i < 8; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1);
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  pdl_wait();                                  // everything below reads what earlier kernels wrote
  if (warp >= 2 && warp < 6) {
    const int t = threadIdx.x - 64;
    if (MODE == MODE_NT) {
      for (int i = t; i < BN; i += kEpiThreads) s_bias[i] = args.bias ? args.bias[i] : 0.f;
      for (int i = t; i < HEADS * BN; i += kEpiThreads) s_head[i] = args.head_w[i];
    }
    if (MODE == MODE_TN) {
      uint32_t* ones = reinterpret_cast<uint32_t*>(s_staging);
      for (int i = t; i < kOnesBytes / 4; i += kEpiThreads) ones[i] = 0x3F803F80u;   // bf16 1.0 x2
      fence_proxy_async_smem();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (MODE != MODE_TN) {
    // =====================================================================================
    // NT / NN: persistent over 128-row tiles
    // =====================================================================================
    if (warp == 0) {
      // ---------------- A producer: activations, HBM -> smem, deep ring ----------------
      if (lane == 0) {
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < args.m_tiles; tile += gridDim.x) {
          for (int kc = 0; kc < args.k_chunks; ++kc) {
            mbar_wait(&empty_a[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_a[s], kABytes);
            tma_load_2d(s_a + s * kABytes, &tmA, &full_a[s], kc * kBlockK, tile * kBlockM);
            if (++s == NA) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 6) {
      // ---------------- B producer: weights, L2 -> smem, shallow ring ----------------
      if (lane == 0) {
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < args.m_tiles; tile += gridDim.x) {
          for (int kc = 0; kc < args.k_chunks; ++kc) {
            mbar_wait(&empty_b[s], ph ^ 1);
            uint8_t* b_s = s_b + s * Cfg::kBBytes;
            mbar_arrive_expect_tx(&full_b[s], Cfg::kBBytes);
            if (MODE == MODE_NT) {
              tma_load_2d(b_s, &tmB, &full_b[s], kc * kBlockK, 0);                      // [BN rows][64 k]
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)                                        // [64 k rows][64 n] boxes
                tma_load_2d(b_s + j * 8192, &tmB, &full_b[s], j * 64, kc * kBlockK);
            }
            if (++s == NB) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer: the whole warp runs the loop converged, one elected lane issues ----------------
      // (a divergent `if (lane == 0)` region makes the compiler wrap every uniform-datapath instruction in a lane loop and
      // rebuild descriptors with 64-bit arithmetic: ~230-320 cycles per MMA instead of the 128-cycle floor)
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, MODE == MODE_NN ? 1 : 0);
      constexpr uint32_t kHiA = desc_hi_sw128(1024), kHiB = desc_hi_sw128(1024);
      constexpr uint32_t kBStep = (MODE == MODE_NT) ? 2u : 128u;                  // 32 B (K-major) or 2048 B (MN-major) per k step
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(s_a), 16);
      const uint32_t b_lo0 = desc_lo_sw128(smem_u32(s_b), MODE == MODE_NT ? 16 : 8192);
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0; int acc = 0; uint32_t acc_ph = 0;
      for (int tile = blockIdx.x; tile < args.m_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < args.k_chunks; ++kc) {
          mbar_wait(&full_b[sb], phb);
          mbar_wait(&full_a[sa], pha);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t al = a_lo0 + sa * (kABytes >> 4), bl = b_lo0 + sb * (Cfg::kBBytes >> 4);
            const int krem = args.k_total - kc * kBlockK;
            umma_bf16(d_tmem, pack64(al, kHiA), pack64(bl, kHiB), idesc, kc != 0);
#pragma unroll
            for (int k = 1; k < 4; ++k)
              if (krem > k * 16) umma_bf16(d_tmem, pack64(al + 2 * k, kHiA), pack64(bl + kBStep * k, kHiB), idesc, 1u);
            umma_commit(&empty_a[sa]);
            umma_commit(&empty_b[sb]);
            if (kc == args.k_chunks - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++sa == NA) { sa = 0; pha ^= 1; }
          if (++sb == NB) { sb = 0; phb ^= 1; }
        }
        acc ^= 1; if (acc == 0) acc_ph ^= 1;
      }
    } else {
      // ---------------- epilogue warps ----------------
      const int q = warp & 3;                       // TMEM lane quarter this warp may access
      const int row = q * 32 + lane;                // row inside the 128-row tile
      const bool issuer = (threadIdx.x == 64);
      int acc = 0; uint32_t acc_ph = 0;
      int sbuf = 0;                                  // staging half buffer written next
      for (int tile = blockIdx.x; tile < args.m_tiles; tile += gridDim.x) {
        const int64_t gr = (int64_t)tile * kBlockM + row;
        const bool row_ok = gr < args.m_rows;
        uint32_t mbits[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) mbits[i] = (MODE == MODE_NN) ? 0xFFFFFFFFu : 0u;
        if (MODE == MODE_NN && args.mask_bits && row_ok) {
          // packed ReLU mask of this row (32 B for BN = 256): issued before the accumulator wait
#pragma unroll
          for (int i = 0; i < NW; ++i) mbits[i] = __ldg(args.mask_bits + gr * NW + i);   // 32 contiguous bytes
        }
        mbar_wait(&tmem_full[acc], acc_ph);
        tcgen05_fence_after();
        const uint32_t t_base = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
        float hacc0 = 0.f, hacc1 = 0.f, hacc2 = 0.f;
        const float relu_lo = (MODE == MODE_NT && args.relu) ? 0.f : -3.0e38f;
        constexpr int kHalves = BN / Cfg::kHalfCols;          // 2 for BN = 256, else 1
        constexpr int kGroupsPerHalf = Cfg::kHalfCols / 32;
#pragma unroll
        for (int h = 0; h < kHalves; ++h) {
          uint8_t* s_half = s_staging + sbuf * Cfg::kHalfBytes;
          // the previous store out of this buffer (two commits ago) must have finished READING it
          if (issuer) tma_store_wait_read1();
          named_bar_sync(1, kEpiThreads);
#pragma unroll
          for (int cg = 0; cg < kGroupsPerHalf; ++cg) {
            const int c = h * kGroupsPerHalf + cg;
            uint32_t v[32];
            tmem_ld_x32(t_base + c * 32, v);
            tmem_ld_wait();
            uint8_t* box = s_half + (cg >> 1) * 16384 + row * 128;
            const uint32_t word = mbits[c];
            uint32_t outbits = 0u;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int lchunk = (cg & 1) * 4 + cc;
              uint4* dst = reinterpret_cast<uint4*>(box + ((lchunk ^ (row & 7)) << 4));
              float x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[cc * 8 + e]);
              if (MODE == MODE_NT) {
                const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c * 32 + cc * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c * 32 + cc * 8 + 4);
                x[0] = fmaxf(x[0] + b0.x, relu_lo); x[1] = fmaxf(x[1] + b0.y, relu_lo);
                x[2] = fmaxf(x[2] + b0.z, relu_lo); x[3] = fmaxf(x[3] + b0.w, relu_lo);
                x[4] = fmaxf(x[4] + b1.x, relu_lo); x[5] = fmaxf(x[5] + b1.y, relu_lo);
                x[6] = fmaxf(x[6] + b1.z, relu_lo); x[7] = fmaxf(x[7] + b1.w, relu_lo);
                if (WMASK) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) outbits |= (x[e] > 0.f) ? (1u << layout::relu_mask_bit(cc * 8 + e)) : 0u;
                }
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = ((word >> layout::relu_mask_bit(cc * 8 + e)) & 1u) ? x[e] : 0.f;
              }
              uint32_t packed[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 p = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
                packed[e] = *reinterpret_cast<uint32_t*>(&p);
              }
              if (MODE == MODE_NT && HEADS > 0) {
                // fused head (model.py:181,194): fp32 dot with the bf16-rounded activation that is stored
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float r0 = __uint_as_float(packed[e] << 16), r1 = __uint_as_float(packed[e] & 0xFFFF0000u);
                  const int j = c * 32 + cc * 8 + 2 * e;
                  hacc0 = fmaf(r0, s_head[j], hacc0); hacc0 = fmaf(r1, s_head[j + 1], hacc0);
                  if (HEADS == 3) {
                    hacc1 = fmaf(r0, s_head[BN + j], hacc1); hacc1 = fmaf(r1, s_head[BN + j + 1], hacc1);
                    hacc2 = fmaf(r0, s_head[2 * BN + j], hacc2); hacc2 = fmaf(r1, s_head[2 * BN + j + 1], hacc2);
                  }
                }
              }
              *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            if (MODE == MODE_NT && WMASK) mbits[c] = outbits;
          }
          if (h == kHalves - 1) {
            // accumulator fully drained -> hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          // staged half tile -> global (TMA store clips rows beyond M); do not wait for it here
          fence_proxy_async_smem();
          named_bar_sync(2, kEpiThreads);
          if (issuer) {
#pragma unroll
            for (int j = 0; j < Cfg::kHalfCols / 64; ++j)
              tma_store_2d(&tmD, s_half + j * 16384, h * Cfg::kHalfCols + j * 64, tile * kBlockM);
            tma_store_commit();
          }
          sbuf ^= 1;
        }
        acc ^= 1; if (acc == 0) acc_ph ^= 1;
        if (MODE == MODE_NT && row_ok) {
          if (WMASK) {
            uint4* mo = reinterpret_cast<uint4*>(args.mask_out + gr * NW);
            if (NW == 8) { mo[0] = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]); mo[1] = make_uint4(mbits[4 % NW], mbits[5 % NW], mbits[6 % NW], mbits[7 % NW]); }
            else {
#pragma unroll
              for (int i = 0; i < NW; ++i) args.mask_out[gr * NW + i] = mbits[i];
            }
          }
          if (HEADS > 0) {
            float* o = args.head_out + gr * 4 + args.head_col;
            o[0] = hacc0 + args.head_b[0];
            if (HEADS == 3) { o[1] = hacc1 + args.head_b[1]; o[2] = hacc2 + args.head_b[2]; }
          }
        }
      }
      if (issuer) tma_store_wait_all0();
    }
  } else {
    // =====================================================================================
    // TN: split-K weight gradient, one 128 x BN output tile per CTA
    // =====================================================================================
    const int m_tile = blockIdx.x % args.m_tiles;
    const int split = blockIdx.x / args.m_tiles;
    const int c0 = split * args.chunks_per_split;
    const int c1 = min(c0 + args.chunks_per_split, args.k_chunks);
    if (warp == 0) {
      if (lane == 0) {
        int s = 0; uint32_t ph = 0;
        const uint64_t pol = l2_policy(args.l2_hints ? 1 : 0);
        for (int kc = c0; kc < c1; ++kc) {
          mbar_wait(&empty_a[s], ph ^ 1);
          uint8_t* a_s = s_a + s * Cfg::kStageBytes;
          uint8_t* b_s = a_s + kABytes;
          mbar_arrive_expect_tx(&full_a[s], Cfg::kStageBytes);
          tma_load_2d_hint(a_s, &tmA, &full_a[s], m_tile * kBlockM, kc * kBlockK, pol);          // [64 pts][64 out]
          tma_load_2d_hint(a_s + 8192, &tmA, &full_a[s], m_tile * kBlockM + 64, kc * kBlockK, pol);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d_hint(b_s + j * 8192, &tmB, &full_a[s], j * 64, kc * kBlockK, pol);
          if (++s == NS) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN > 256 ? 256 : BN, 1, 1);
      constexpr uint32_t idesc2 = make_idesc_bf16(kBlockM, BN > 256 ? BN - 256 : 16, 1, 1);
      constexpr uint32_t idesc1 = make_idesc_bf16(kBlockM, 16, 1, 1);
      constexpr uint32_t kHi = desc_hi_sw128(1024);
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(s_a), 8192);
      const uint32_t o_lo = desc_lo_sw128(smem_u32(s_staging), 8192);
      int s = 0; uint32_t ph = 0;
      for (int kc = c0; kc < c1; ++kc) {
        mbar_wait(&full_a[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t al = a_lo0 + s * (Cfg::kStageBytes >> 4), bl = al + (kABytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (k != 0) ? 1u : (uint32_t)(kc != c0);
            const uint64_t adesc = pack64(al + k * 128, kHi);
            umma_bf16(tmem_base, adesc, pack64(bl + k * 128, kHi), idesc, accum);
            if (BN > 256)                                                 // columns 256.. : a second, narrower MMA
              umma_bf16(tmem_base + 256, adesc, pack64(bl + 4 * 512 + k * 128, kHi), idesc2, accum);
            umma_bf16(tmem_base + BN, adesc, pack64(o_lo, kHi), idesc1, accum);      // column sums of A (bias gradient)
          }
          umma_commit(&empty_a[s]);
          if (kc == c1 - 1) umma_commit(&tmem_full[0]);
        }
        __syncwarp();
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    } else if (warp < 6) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      float* out = args.partial + (size_t)blockIdx.x * (kBlockM * (BN + 1));
      if (c1 > c0) {
        mbar_wait(&tmem_full[0], 0);
        tcgen05_fence_after();
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_base + c * 32, v);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(out + (size_t)row * BN + c * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                 __uint_as_float(v[4 * e + 3]));
        }
        uint32_t v16[16];
        tmem_ld_x16(t_base + BN, v16);
        tmem_ld_wait();
        out[(size_t)kBlockM * BN + row] = __uint_as_float(v16[0]);
      } else {
        for (int c = 0; c < BN; ++c) out[(size_t)row * BN + c] = 0.f;
        out[(size_t)kBlockM * BN + row] = 0.f;
      }
      tcgen05_fence_before();
    }
  }

  // ---------------- teardown ----------------
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Deterministic split-K reduction + scatter into the (unpadded) parameter-gradient layout.
// Output element o (row-major over [nrows][ncols], followed by nrows column sums) is summed by 8 threads
// (each takes every 8th split), then combined in a fixed order -> bit-reproducible, and 8x the loads in flight
// of a thread-per-element loop (the partial tiles total up to 19 MB per layer).
constexpr int kRedParts = 8;
// number of 32-item blocks of one reduce descriptor: float4 items over the 4-column-aligned cover of [col0, col0 + ncols)
// of every row, then one scalar item per row for the column sums
__host__ __device__ inline int splitk_vec_cols(int col0, int ncols) { return ((col0 + ncols + 3) >> 2) - (col0 >> 2); }
__host__ __device__ inline int splitk_reduce_blocks(int nrows, int col0, int ncols) {
  return (nrows * splitk_vec_cols(col0, ncols) + nrows + 31) / 32;
}
// Each item is summed by 8 threads (every 8th split each, 16-byte loads: four columns at a time -- the scalar version ran
// at 3.1 TB/s on the 390 MB of partials of a step) and combined in a fixed order -> bit-reproducible.
__device__ __forceinline__ void splitk_reduce_block(const float* __restrict__ partial, int m_tiles, int splits, int BN, int row0,
                                                    int nrows, int col0, int ncols, float* __restrict__ dst, int64_t dst_ld,
                                                    float* __restrict__ colsum_dst, int block) {
  __shared__ float4 s_part[kRedParts][32];
  const int e = threadIdx.x & 31;            // item within the CTA's 32-item block
  const int part = threadIdx.x >> 5;         // which splits this thread sums
  const int ncv = splitk_vec_cols(col0, ncols), cv0 = col0 >> 2;
  const int total_v = nrows * ncv;
  const int o = block * 32 + e;
  const size_t blk = (size_t)kBlockM * (BN + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  bool live = false, vec = false;
  size_t off = 0;
  int mt = 0, r_out = 0, c_first = 0;
  if (o < total_v) {
    live = dst != nullptr; vec = true;
    r_out = o / ncv;
    const int r = row0 + r_out;
    c_first = (cv0 + o % ncv) << 2;
    mt = r / kBlockM;
    off = (size_t)(r % kBlockM) * BN + c_first;
  } else if (o < total_v + nrows) {
    live = colsum_dst != nullptr;
    r_out = o - total_v;
    const int r = row0 + r_out;
    mt = r / kBlockM;
    off = (size_t)kBlockM * BN + (r % kBlockM);
  }
  if (live) {
    if (vec) {
#pragma unroll 4
      for (int s = part; s < splits; s += kRedParts) {
        const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)s * m_tiles + mt) * blk + off);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    } else {
#pragma unroll 4
      for (int s = part; s < splits; s += kRedParts) acc.x += partial[((size_t)s * m_tiles + mt) * blk + off];
    }
  }
  s_part[part][e] = acc;
  __syncthreads();
  if (part == 0 && live) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kRedParts; ++p) { const float4 v = s_part[p][e]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    if (vec) {
      const float tv[4] = {t.x, t.y, t.z, t.w};
      float* drow = dst + (int64_t)r_out * dst_ld;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c_first + k;
        if (c >= col0 && c < col0 + ncols) drow[c - col0] = tv[k];
      }
    } else {
      colsum_dst[r_out] = t.x;
    }
  }
}
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int m_tiles, int splits, int BN, int row0, int nrows, int col0,
                     int ncols, float* __restrict__ dst, int64_t dst_ld, float* __restrict__ colsum_dst) {
  splitk_reduce_block(partial, m_tiles, splits, BN, row0, nrows, col0, ncols, dst, dst_ld, colsum_dst, blockIdx.x);
}
__global__ void __launch_bounds__(256) splitk_reduce_batch_kernel(const __grid_constant__ TnBatch b) {
  pdl_launch_dependents();
  pdl_wait();
  int i = 0;
  while (i + 1 < b.n && (int)blockIdx.x >= b.d[i + 1].block0) ++i;
  const TnReduceDesc& d = b.d[i];
  splitk_reduce_block(d.partial, d.m_tiles, d.splits, d.BN, d.row0, d.nrows, d.col0, d.ncols, d.dst, d.dst_ld, d.colsum_dst,
                      (int)blockIdx.x - d.block0);
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor [outer][inner] with leading dimension ld (elements), box [box_outer][64], SWIZZLE_128B
int make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return RN_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8) != 0 || inner == 0 || outer == 0) return RN_ERR_INVALID_ARG;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RN_OK : RN_ERR_DRIVER;
}

// ---- optional per-launch CUDA-event timing (bench.py roofline: live kernel durations on the
// launching stream inside the timed region) ----
struct ProfRec { cudaEvent_t a, b; int mode; double flops; int launches; };
static bool g_prof_on = false;
int g_prof_suppress = 0;     // > 0: inside a span that is recorded as a whole (the overlapped backward): skip per-launch records
static std::vector<ProfRec> g_prof;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
double g_prof_next_flops = 0.0;
int g_sm_limit_dgrad = 0, g_sm_limit_wgrad = 0;    // rn_set_flag(7 / 8, n): experiments -- run that kernel family on n SMs
int g_l2_hints = 0;          // rn_set_flag(6, v): bit 0 = chain kernels, bit 1 = split-K weight-gradient kernels, bit 3 = NO policies on the chain -> stream hand-off

void prof_begin(int mode, cudaStream_t st, int* slot) {
  *slot = -1;
  if (!g_prof_on || g_prof.size() >= 16384 || (g_prof_suppress > 0 && mode != 3)) return;
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (!g_prof_pool.empty()) { ev = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) return;
  cudaEventRecord(ev.first, st);
  g_prof.push_back({ev.first, ev.second, mode, g_prof_next_flops, mode == 3 ? 2 : 1});
  *slot = (int)g_prof.size() - 1;
}
void prof_end(int slot, cudaStream_t st) {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}

template <int BN, int MODE, int HEADS = 0, bool WMASK = false>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tD, const GemmArgs& args, int grid,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN, MODE>;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    RN_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BN, MODE, HEADS, WMASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
  int slot;
  prof_begin(MODE, st, &slot);
  RN_CUDA_CHECK(launch_maybe_pdl(gemm_kernel<BN, MODE, HEADS, WMASK>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, st, tA, tB, tD, args));
  prof_end(slot, st);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int check_arch() {
  static int cached[kMaxDevices];            // 0 = unknown, 1 = ok, 2 = unsupported (per device)
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return RN_ERR_CUDA;
  int& c = cached[dev & (kMaxDevices - 1)];
  if (c == 0) {
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return RN_ERR_CUDA;
    c = (major == 10) ? 1 : 2;
  }
  return c == 1 ? RN_OK : RN_ERR_UNSUPPORTED_ARCH;
}

// D[M,N] = act(A[M,K] B[N,K]^T + bias)        (forward layer)
int gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int64_t M, int N, int K,
            const float* bias, int relu, cudaStream_t st, uint32_t* mask_out, int n_heads, const float* head_w,
            const float* head_b, float* head_out, int head_col) {
  int rc = check_arch();
  if (rc != RN_OK) return rc;
  RN_REQUIRE(M > 0 && K > 0 && K % 8 == 0 && (N == 256 || N == 128 || N == 64));
  CUtensorMap tA, tB, tD;
  if ((rc = make_tmap(&tA, A, K, M, lda, kBlockM)) != RN_OK) return rc;
  if ((rc = make_tmap(&tB, B, K, N, ldb, N)) != RN_OK) return rc;
  if ((rc = make_tmap(&tD, D, N, M, ldd, kBlockM)) != RN_OK) return rc;
  GemmArgs a{};
  a.m_tiles = (int)ceil_div(M, kBlockM); a.k_chunks = (int)ceil_div(K, kBlockK); a.k_total = K; a.bias = bias; a.relu = relu;
  RN_REQUIRE(n_heads == 0 || ((n_heads == 1 || n_heads == 3) && head_w && head_b && head_out));
  a.mask_out = mask_out;
  a.m_rows = M; a.n_heads = n_heads; a.head_w = head_w; a.head_b = head_b; a.head_out = head_out; a.head_col = head_col;
  const int grid = a.m_tiles < num_sms() ? a.m_tiles : num_sms();
  g_prof_next_flops = 2.0 * (double)M * N * K;
  if (N == 256) {
    RN_REQUIRE(n_heads <= 1);
    if (n_heads == 1) return mask_out ? launch_gemm<256, MODE_NT, 1, true>(tA, tB, tD, a, grid, st)
                                      : launch_gemm<256, MODE_NT, 1, false>(tA, tB, tD, a, grid, st);
    return mask_out ? launch_gemm<256, MODE_NT, 0, true>(tA, tB, tD, a, grid, st)
                    : launch_gemm<256, MODE_NT, 0, false>(tA, tB, tD, a, grid, st);
  }
  RN_REQUIRE(mask_out == nullptr);
  if (N == 128) {
    RN_REQUIRE(n_heads == 0 || n_heads == 3);
    return n_heads == 3 ? launch_gemm<128, MODE_NT, 3, false>(tA, tB, tD, a, grid, st)
                        : launch_gemm<128, MODE_NT, 0, false>(tA, tB, tD, a, grid, st);
  }
  RN_REQUIRE(n_heads == 0);
  return launch_gemm<64, MODE_NT>(tA, tB, tD, a, grid, st);
}

// D[M,N] = (A[M,K] B[K,N]) .* mask   (data gradient; mask_bits = packed ReLU mask [M][N/32] or null)
int gemm_nn(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int64_t M, int N, int K,
            const uint32_t* mask_bits, cudaStream_t st) {
  int rc = check_arch();
  if (rc != RN_OK) return rc;
  RN_REQUIRE(M > 0 && K > 0 && K % 8 == 0 && (N == 256 || N == 64));
  CUtensorMap tA, tB, tD;
  if ((rc = make_tmap(&tA, A, K, M, lda, kBlockM)) != RN_OK) return rc;
  if ((rc = make_tmap(&tB, B, N, K, ldb, 64)) != RN_OK) return rc;
  if ((rc = make_tmap(&tD, D, N, M, ldd, kBlockM)) != RN_OK) return rc;
  GemmArgs a{};
  a.m_tiles = (int)ceil_div(M, kBlockM); a.k_chunks = (int)ceil_div(K, kBlockK); a.k_total = K;
  a.m_rows = M; a.mask_bits = mask_bits;
  const int grid = a.m_tiles < num_sms() ? a.m_tiles : num_sms();
  g_prof_next_flops = 2.0 * (double)M * N * K;
  if (N == 256) return launch_gemm<256, MODE_NN>(tA, tB, tD, a, grid, st);
  return launch_gemm<64, MODE_NN>(tA, tB, tD, a, grid, st);
}

size_t gemm_tn_scratch_bytes() { return (size_t)(num_sms() + 8) * kBlockM * 321 * sizeof(float); }

// Weight / bias gradient: partial[split][m_tile] = A[Kslice, Mo]^T B[Kslice, N] (+ column sums of A),
// then gemm_tn_reduce sums the splits in a fixed order and scatters rows [row0,row0+nrows) x first
// ncols columns into dst (leading dimension dst_ld) and the column sums into colsum_dst.
int gemm_tn_launch(const void* A, int64_t lda, int Mo, const void* B, int64_t ldb, int N, int64_t K, float* scratch,
                   size_t scratch_bytes, TnInfo* info, cudaStream_t st) {
  int rc = check_arch();
  if (rc != RN_OK) return rc;
  RN_REQUIRE(K > 0 && Mo > 0 && (N == 320 || N == 256 || N == 64) && scratch && info);
  CUtensorMap tA, tB;
  if ((rc = make_tmap(&tA, A, Mo, K, lda, 64)) != RN_OK) return rc;
  if ((rc = make_tmap(&tB, B, N, K, ldb, 64)) != RN_OK) return rc;
  GemmArgs a{};
  a.m_tiles = (int)ceil_div(Mo, kBlockM);
  a.k_chunks = (int)ceil_div(K, kBlockK);
  a.k_total = (int)(K > 0x7fffffff ? 0x7fffffff : K);
  const int sms = (g_sm_limit_wgrad > 0 && g_sm_limit_wgrad < num_sms()) ? g_sm_limit_wgrad : num_sms();
  int splits = sms / a.m_tiles;
  if (splits > a.k_chunks) splits = a.k_chunks;
  if (splits < 1) splits = 1;
  a.chunks_per_split = (int)ceil_div(a.k_chunks, splits);
  a.splits = (int)ceil_div(a.k_chunks, a.chunks_per_split);
  a.partial = scratch;
  a.l2_hints = (g_l2_hints >> 1) & 1;
  RN_REQUIRE((size_t)a.splits * a.m_tiles * kBlockM * (N + 1) * sizeof(float) <= scratch_bytes);
  const int grid = a.m_tiles * a.splits;
  g_prof_next_flops = 2.0 * (double)K * N * Mo;
  if (N == 320) rc = launch_gemm<320, MODE_TN>(tA, tB, tA, a, grid, st);
  else if (N == 256) rc = launch_gemm<256, MODE_TN>(tA, tB, tA, a, grid, st);
  else rc = launch_gemm<64, MODE_TN>(tA, tB, tA, a, grid, st);
  info->m_tiles = a.m_tiles; info->splits = a.splits; info->N = N; info->scratch = scratch;
  return rc;
}

int gemm_tn_reduce(const TnInfo& info, int row0, int nrows, int col0, int ncols, float* dst, int64_t dst_ld,
                   float* colsum_dst, cudaStream_t st) {
  RN_REQUIRE(row0 >= 0 && nrows > 0 && col0 >= 0 && ncols > 0 && col0 + ncols <= info.N &&
             row0 + nrows <= info.m_tiles * kBlockM);
  splitk_reduce_kernel<<<splitk_reduce_blocks(nrows, col0, ncols), 256, 0, st>>>(info.scratch, info.m_tiles, info.splits, info.N, row0, nrows,
                                                           col0, ncols, dst, dst_ld, colsum_dst);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int tn_batch_add(TnBatch* b, const TnInfo& info, int row0, int nrows, int col0, int ncols, float* dst, int64_t dst_ld,
                 float* colsum_dst) {
  RN_REQUIRE(b && b->n < kTnBatchMax && row0 >= 0 && nrows > 0 && col0 >= 0 && ncols > 0 && col0 + ncols <= info.N &&
             row0 + nrows <= info.m_tiles * kBlockM);
  TnReduceDesc& d = b->d[b->n++];
  d.partial = info.scratch; d.dst = dst; d.colsum_dst = colsum_dst; d.dst_ld = dst_ld;
  d.m_tiles = info.m_tiles; d.splits = info.splits; d.BN = info.N; d.row0 = row0; d.nrows = nrows; d.col0 = col0; d.ncols = ncols;
  d.block0 = b->total_blocks;
  b->total_blocks += splitk_reduce_blocks(nrows, col0, ncols);
  return RN_OK;
}
int gemm_tn_reduce_batch(const TnBatch& b, cudaStream_t st) {
  if (b.n == 0) return RN_OK;
  RN_CUDA_CHECK(launch_maybe_pdl(splitk_reduce_batch_kernel, dim3(b.total_blocks), dim3(256), 0, st, b));
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

using namespace rn;

extern "C" {

size_t rn_gemm_scratch_bytes(void) { return gemm_tn_scratch_bytes(); }

int rn_prof_enable(int on) {
  g_prof_on = on != 0;
  return RN_OK;
}

// Synchronises the device, then sums the recorded GEMM launches by mode (0 NT, 1 NN, 2 TN, 3 = the NN chain and the
// TN stream running side by side, timed as one span): total milliseconds, executed (padded) FLOPs and launch counts;
// clears the record.  The arrays hold FOUR entries.
int rn_prof_collect(double* ms4_host, double* flops4_host, int* launches4_host) {
  RN_REQUIRE(ms4_host && flops4_host && launches4_host);
  RN_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < 4; ++i) { ms4_host[i] = 0.0; flops4_host[i] = 0.0; launches4_host[i] = 0; }
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms4_host[r.mode] += ms; flops4_host[r.mode] += r.flops; launches4_host[r.mode] += r.launches;
    }
    g_prof_pool.push_back({r.a, r.b});
  }
  g_prof.clear();
  return RN_OK;
}

int rn_gemm_bf16(int mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int64_t M, int N,
                 int64_t K, const float* bias, int relu, const uint32_t* mask_bits, uint32_t* mask_out, float* colsum_out,
                 void* scratch, size_t scratch_bytes, rn_stream_t stream) {
  RN_REQUIRE(A && B && D);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == MODE_NT)
    return gemm_nt(A, lda, B, ldb, D, ldd, M, N, (int)K, bias, relu, st, mask_out, 0, nullptr, nullptr, nullptr, 0);
  if (mode == MODE_NN) return gemm_nn(A, lda, B, ldb, D, ldd, M, N, (int)K, mask_bits, st);
  if (mode == MODE_TN) {
    TnInfo info;
    int rc = gemm_tn_launch(A, lda, (int)M, B, ldb, N, K, (float*)scratch, scratch_bytes, &info, st);
    if (rc != RN_OK) return rc;
    return gemm_tn_reduce(info, 0, (int)M, 0, N, (float*)D, ldd, colsum_out, st);
  }
  return RN_ERR_INVALID_ARG;
}

}  // extern "C"
