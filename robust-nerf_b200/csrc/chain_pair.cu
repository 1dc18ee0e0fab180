// chain_pair.cu -- the NeRF MLP forward (noisy_src/model.py:145-196) as ONE persistent launch per network in which the
// activations never leave the SM between layers.
//
// A pair of CTAs (cluster of 2 on one TPC) pushes two "pair tiles" of 256 points (128 rows per CTA) through all ten GEMM
// layers.  Per CTA and tile slot a 64 KiB shared-memory buffer holds the current 128 x 256 bf16 activation in the
// SWIZZLE_128B K-major layout the tensor core reads: layer l's epilogue converts the fp32 accumulator (TMEM) to bf16
// and writes it IN PLACE over layer l-1's activation (whose MMAs have completed), and layer l+1's MMAs read it straight
// from there -- no TMA load, no L2 round trip, no staging copy.  In training the same buffer is also the source of the
// TMA store that keeps the activation for the backward pass; in inference nothing but raw[M,4] is written at all.
// The skip concat [x_enc | h] and the view concat [features | d_enc] are extra 64-wide K chunks in a 16 KiB side buffer
// (TMA-loaded).  MMAs are tcgen05.mma.cta_group::2 (M = 256 per pair): each CTA streams only HALF of every weight chunk
// (16 KiB instead of 32), which halves the weight traffic through shared memory and makes a 4-deep ring fit.
// The two tile slots ping-pong: slot X's MMAs run while slot Y's epilogue drains, so the tensor pipe is busy as long as
// an epilogue (8 warps) is not slower than a layer's MMAs.
//
// Shared-memory traffic per 128 x 256 x 256 tile-layer and SM: 64 KiB weight writes + 64 KiB weight reads + 64 KiB
// activation reads + 64 KiB activation writes (+ 64 KiB store reads in training) = 2048-2560 cycles at 128 B/clk, i.e.
// at the 2048-cycle MMA floor -- the per-layer kernel moves 512 KiB (gemm_tcgen05.cu, profiles/r01_chain_experiments.md).
//
// Warp roles (608 threads): 0 = TMA producer (weights + side chunks), 1 = MMA issuer (pair leader only) + TMEM owner,
// 2..17 = epilogue (TMEM lane quarter = warp % 4, column quarter = (warp - 2) / 4), 18 = TMA store warp (training).
// The epilogue is the pacing stage (profiles/r01_pair_experiments.md): sixteen warps (four per scheduler) hide the
// TMEM-load and shared-store latencies, and the arithmetic is packed two columns per instruction (FADD2 for the bias,
// F2FP to bf16x2, HMNMX2 for the ReLU; the mask bits come from the packed word with an add, a shift and a LOP3).
#include "common.cuh"
#include "ptx.cuh"
#include "gemm.h"
#include "mlp_layout.h"
#include "pe.cuh"
#include <cuda_bf16.h>
#include <mutex>

namespace rn {

using namespace ptx;

int make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer);
void prof_begin(int mode, cudaStream_t st, int* slot);
void prof_end(int slot, cudaStream_t st);
extern double g_prof_next_flops;
extern int g_l2_hints, g_sm_limit_dgrad, g_ws_stagger_us;

constexpr int kPairThreads = 608;         // 19 warps
constexpr int kPairThreadsPE = 640;       // + one warp: with the (idle in inference) store warp, two encoder warps, one per tile slot
constexpr int kPairMaxLayers = 12;
constexpr int kPairMaxChunks = 8;
constexpr int kChunkBytes = 16384;              // [128 rows][64 bf16], SWIZZLE_128B
constexpr int kActBytes = 4 * kChunkBytes;      // one 128 x 256 activation
constexpr int kAuxBytes = kChunkBytes;
constexpr int kBStageBytes = 16384;             // half of a [256][64] weight chunk
constexpr int kNBStages = 4;
constexpr int kPairMisc = 2048;                 // head partial sums [3][128] fp32 + barriers
constexpr int kPairSmem = 2 * kActBytes + 2 * kAuxBytes + kNBStages * kBStageBytes + kPairMisc + 1024 /*alignment*/;
static_assert(kPairSmem <= 232448, "pair chain kernel exceeds the 227 KiB shared-memory limit");

struct PairLayer {
  int n;                       // output columns: 256 or 128
  int k_chunks;                // 64-wide K chunks of the A operand
  int8_t a_src[kPairMaxChunks];   // per chunk: 0..3 = activation chunk, 4 = side buffer
  int aux_load;                // 0, or 1 + index of the side tensor map loaded before this layer
  int aux_release;             // the side buffer is free again once this layer's MMAs completed
  int relu, heads, head_col;
  int bias_off, head_w_off, head_b_off;   // float offsets into the fp32 constants
  int store;                   // training: TMA-store the output tile through tmD
  uint32_t* mask_out;          // packed ReLU mask [M][8] words, or null
};
// measurement knobs, compiled in only with -DRN_EXPERIMENTS (results are WRONG when any bit is set):
// bit 0: no weight loads   bit 1: no TMA stores   bit 2: no side-chunk loads   bit 3: epilogue reads TMEM but skips the math
// bit 4: epilogue does not even read TMEM   bit 7: no fence.proxy.async after the epilogue's shared-memory writes
#ifdef RN_EXPERIMENTS
extern int g_chain_dbg;
#define RN_PDBG(p, bit) ((p).dbg & (bit))
#else
#define RN_PDBG(p, bit) (0)
#endif
#ifdef RN_EXPERIMENTS
// timeline of cluster 0 (scripts/pair_timeline.py): role r of CTA rank c appends (tag, clock64, globaltimer) triples
constexpr int kTlRoles = 4, kTlEntries = 2048;
long long* g_pair_timeline = nullptr;        // [2 ranks][kTlRoles][kTlEntries][3]
long long* g_pair_timeline_bwd = nullptr;    // same, data-gradient chain
struct Tl {
  long long* base; int n;
  __device__ __forceinline__ void init(long long* buf, int rank, int role, bool on) {
    base = (on && buf) ? buf + ((size_t)(rank * kTlRoles + role) * kTlEntries) * 3 : nullptr; n = 0;
  }
  __device__ __forceinline__ void rec(int tag) {
    if (base && n < kTlEntries) {
      unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      base[n * 3] = tag; base[n * 3 + 1] = clock64(); base[n * 3 + 2] = (long long)g; ++n;
    }
  }
};
#define RN_TL_DECL(name, role, cond) Tl name; name.init(p.timeline, (int)rank, role, (cond) && cluster_id == 0)
#define RN_TL(name, tag) name.rec(tag)
#else
#define RN_TL_DECL(name, role, cond)
#define RN_TL(name, tag)
#endif
struct PairParams {
  long long* timeline;
  int dbg;
  CUtensorMap tmB[kPairMaxLayers], tmD[kPairMaxLayers], tmAux[2];
  PairLayer L[kPairMaxLayers];
  int n_layers, n_ptiles;      // pair tiles of 256 rows
  int64_t m_rows;
  int const_slot;              // which slot of c_pair_consts2 holds this network's fp32 section
  int l2_hints;                // rn_set_flag(6): weights evict_last, streamed activations evict_first
  float* raw;
  // PE-fused inference (mlp_chain_pair_kernel<false, true>): the kernel encodes the points itself and takes the view
  // branch's direction term as a per-ray bias of the last layer
  const float* pts;            // [M][3]
  const float* dirvec;         // [M / group][128]  (encode.cu: dir_bias_kernel)
  int64_t group;               // consecutive points per ray
};

// Biases and fp32 head weights of the network being evaluated (layout::kF32Elems floats), copied device-to-device on the
// launching stream before each launch.  Every lane of an epilogue warp needs the SAME values: read through the
// constant cache they cost no LSU/L1 cycles (a uniform LDG.128 per 4 columns made the load/store unit -- shared with the
// tensor core's operand reads -- the bottleneck of the epilogue, profiles/r01_pair_experiments.md; reading the head
// weights with warp-uniform __ldg instead cost +10 % of the training forward, profiles/r02_ab_log.md).
//
// The staging area is NOT one global: it is kConstSlots slots keyed by the packed-weight buffer (acquire_const_slot
// below), so forwards of different networks -- coarse and fine overlapped on two streams, an evaluation render beside a
// training step, several host threads -- never share a slot, and two launches of the SAME network write identical
// bytes.  A fifth distinct network evicts the least recently used slot after a device synchronisation.
constexpr int kConstSlots = 4;
constexpr int kConstSlotFloat2 = 1664 + 128;      // + slack: the bias staging copies 256 floats from any bias offset
__constant__ float2 c_pair_consts2[kConstSlots][kConstSlotFloat2];
static_assert(layout::kF32Elems <= 2 * 1664, "constant staging slot too small");
static_assert(sizeof(float2) * kConstSlots * kConstSlotFloat2 <= 60 * 1024, "constant bank exceeded");

__device__ __forceinline__ uint64_t fadd2(uint32_t a_lo, uint32_t a_hi, float2 b) {
  uint64_t a, bb, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a_lo), "r"(a_hi));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(bb));
  return r;
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ua, ub, uc, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(ua), "l"(ub), "l"(uc));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}

// One 8-column chunk of a row: bias (+ReLU) -> bf16 -> 16 bytes of the in-place activation tile (+ mask bits, heads).
// c = chunk index inside the layer output (columns 8c .. 8c+7).
template <int HEADS, bool RELU, bool WMASK>
__device__ __forceinline__ void pair_epilogue_chunk(const uint32_t (&v)[8], int c, uint8_t* s_tile, int row, const float* s_bias,
                                                    const float2* consts2, int head_w2_off, int n, uint32_t& outbits,
                                                    float2& hh0, float2& hh1, float2& hh2) {
  const int G = c >> 2, cc = c & 3;               // 32-column group and chunk inside it
  const float4 bA = *reinterpret_cast<const float4*>(s_bias + c * 8);          // broadcast LDS.128 from the per-layer staging
  const float4 bB = *reinterpret_cast<const float4*>(s_bias + c * 8 + 4);
  const float2 bsel[4] = {make_float2(bA.x, bA.y), make_float2(bA.z, bA.w), make_float2(bB.x, bB.y), make_float2(bB.z, bB.w)};
  uint32_t packed[4], fl[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint64_t r = fadd2(v[2 * e], v[2 * e + 1], bsel[e]);
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
    if (RELU) {
      // convert and ReLU in one instruction (F2FP.RELU): round(max(x, 0)) == max(round(x), 0)
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(packed[e]) : "f"(hi), "f"(lo));
    } else {
      __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
      packed[e] = *reinterpret_cast<uint32_t*>(&pk);
    }
    // after the ReLU both halves are +0 or positive bit patterns: an unsigned 16-bit min with 1 is the "> 0" flag of
    // each half (bits 0 and 16), one instruction per pair
    if (WMASK) fl[e] = __vminu2(packed[e], 0x00010001u);
    if (HEADS > 0) {
      // fused head (model.py:181,194): fp32 dot with the bf16-rounded activation; even and odd columns accumulate in
      // the two halves of a packed fp32 pair
      const int w2 = head_w2_off + c * 4 + e;
      const float2 a = make_float2(__uint_as_float(packed[e] << 16), __uint_as_float(packed[e] & 0xFFFF0000u));
      hh0 = ffma2(a, consts2[w2], hh0);
      if (HEADS == 3) {
        hh1 = ffma2(a, consts2[w2 + (n >> 1)], hh1);
        hh2 = ffma2(a, consts2[w2 + n], hh2);
      }
    }
  }
  // the four pairs of this chunk are pairs 4cc .. 4cc+3 of the group: flags to bits 4cc + e and 16 + 4cc + e
  if (WMASK) outbits |= (fl[0] + 2u * fl[1] + 4u * fl[2] + 8u * fl[3]) << (cc * 4);
  uint8_t* box = s_tile + (G >> 1) * kChunkBytes + row * 128;
  const int lchunk = (G & 1) * 4 + cc;
  *reinterpret_cast<uint4*>(box + ((lchunk ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

// NG groups of 32 accumulator columns of one row, starting at group g0: TMEM -> ... -> in-place activation tile.
// A ROLLED loop over 8-column chunks (two per iteration, TMEM loads double-buffered): the fully unrolled version was
// ~1,000 straight-line instructions per layer type and the epilogue warps spent 40 % of their stall samples waiting for
// instruction fetch (stall_no_inst, profiles/r01_ncu_full_prof_fwd_pair.raw.csv); this body is ~90 instructions that
// every warp re-executes out of the instruction cache.  g0 stays a run-time value so that all sixteen epilogue warps
// share one instruction stream.
template <int NG, int HEADS, bool RELU, bool WMASK>
__device__ __forceinline__ void pair_epilogue(uint32_t t_addr, uint8_t* s_tile, int row, int g0, const float* s_bias,
                                              const float2* consts2, int head_w2_off, int n, uint32_t (&mb)[2],
                                              float& h0, float& h1, float& h2) {
  constexpr int NC = NG * 4;                      // 8-column chunks handled by this warp
  const int c0 = g0 * 4;
  float2 hh0 = make_float2(0.f, 0.f), hh1 = make_float2(0.f, 0.f), hh2 = make_float2(0.f, 0.f);
  uint32_t va[16], vb[16];
  uint32_t outbits = 0u;
  auto two_chunks = [&](const uint32_t (&v)[16], int c) {
    uint32_t lo8[8], hi8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { lo8[k] = v[k]; hi8[k] = v[8 + k]; }
    pair_epilogue_chunk<HEADS, RELU, WMASK>(lo8, c, s_tile, row, s_bias, consts2, head_w2_off, n, outbits, hh0, hh1, hh2);
    pair_epilogue_chunk<HEADS, RELU, WMASK>(hi8, c + 1, s_tile, row, s_bias, consts2, head_w2_off, n, outbits, hh0, hh1, hh2);
  };
  tmem_ld_x16(t_addr, va);
#pragma unroll 1
  for (int i = 0; i < NC; i += 4) {               // one 32-column group per iteration, two 16-column TMEM loads in flight
    tmem_ld_wait();
    tmem_ld_x16(t_addr + (i + 2) * 8, vb);
    two_chunks(va, c0 + i);
    tmem_ld_wait();
    if (i + 4 < NC) tmem_ld_x16(t_addr + (i + 4) * 8, va);
    two_chunks(vb, c0 + i + 2);
    if (i < 4) mb[0] = outbits; else mb[1] = outbits;
    outbits = 0u;
  }
  if (HEADS > 0) { h0 = hh0.x + hh0.y; h1 = hh1.x + hh1.y; h2 = hh2.x + hh2.y; }
}

template <bool TRAIN, bool PE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PE ? kPairThreadsPE : kPairThreads, 1)
mlp_chain_pair_kernel(const __grid_constant__ PairParams p) {
  static_assert(!(TRAIN && PE), "the in-kernel encoder is built for inference (training stores x_enc / d_enc for the weight gradients)");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_act = smem;                                   // [2 slots][4 chunks][16 KiB]
  uint8_t* s_aux = smem + 2 * kActBytes;                   // [2 slots][16 KiB]
  uint8_t* s_b = s_aux + 2 * kAuxBytes;                    // [4 stages][16 KiB]
  float* s_hx = reinterpret_cast<float*>(s_b + kNBStages * kBStageBytes);   // [3][128] head partials of column half 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_hx + 3 * 128);
  uint64_t* full_b = bars;               // [4]  leader: weight chunk landed in BOTH CTAs
  uint64_t* empty_b = bars + 4;          // [4]  each CTA: stage consumed (multicast commit)
  uint64_t* aux_full = bars + 8;         // [2]  leader
  uint64_t* aux_empty = bars + 10;       // [2]  each CTA (multicast commit)
  uint64_t* act_ready = bars + 12;       // [2]  leader: 32 epilogue warps (both CTAs) wrote the tile and drained TMEM
  uint64_t* tmem_full = bars + 14;       // [2]  each CTA (multicast commit)
  uint64_t* staged = bars + 16;          // [2]  epilogue -> store warp (training)
  uint64_t* store_done = bars + 18;      // [2]  store warp -> epilogue: the tile has been read out, overwrite allowed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  // per-layer bias staging [256] fp32 in the last KiB of the allocation -- the alignment slack, which is free because the
  // dynamic shared memory of a kernel without static shared memory starts 1 KiB-aligned (checked: trap otherwise)
  float* s_bias = reinterpret_cast<float*>(smem + 2 * kActBytes + 2 * kAuxBytes + kNBStages * kBStageBytes + kPairMisc);
  if (smem != smem_raw) {
    if (threadIdx.x == 0) printf("rnerf_b200: dynamic shared memory is not 1 KiB aligned (block %d)\n", (int)blockIdx.x);
    __trap();
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    for (int l = 0; l < p.n_layers; ++l) { prefetch_tmap(&p.tmB[l]); if (TRAIN) prefetch_tmap(&p.tmD[l]); }
    prefetch_tmap(&p.tmAux[0]); prefetch_tmap(&p.tmAux[1]);
    for (int i = 0; i < kNBStages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&aux_full[i], PE ? 2 : 1); mbar_init(&aux_empty[i], 1);      // PE: one arrive per CTA's encoder warp
      mbar_init(&act_ready[i], 32); mbar_init(&tmem_full[i], 1);
      mbar_init(&staged[i], 16); mbar_init(&store_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                    // both CTAs' barriers exist before any remote arrive / multicast commit
  tcgen05_fence_after();
  pdl_wait();                            // PDL: everything below reads what earlier kernels wrote
  const uint32_t tmem_base = *tmem_slot;
  const int n_groups = (p.n_ptiles + 1) >> 1;

  if (warp == 0) {
    // ---------------- TMA producer: this CTA's half of every weight chunk, and its rows of the side chunks ----------------
    const uint32_t full_b_leader = mapa_u32(smem_u32(full_b), 0);
    const uint32_t aux_full_leader = mapa_u32(smem_u32(aux_full), 0);
    int s = 0; uint32_t ph = 0;
    uint32_t aux_n0 = 0, aux_n1 = 0;
    const uint64_t pol_w = l2_policy(p.l2_hints ? 2 : 0), pol_s = l2_policy(p.l2_hints ? 1 : 0);
    RN_TL_DECL(tl, 3, lane == 0);
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      const int tiles_here = min(2, p.n_ptiles - grp * 2);
      for (int l = 0; l < p.n_layers; ++l) {
        const int n = p.L[l].n, k_chunks = p.L[l].k_chunks, aux_load = p.L[l].aux_load;
        for (int slot = 0; slot < tiles_here; ++slot) {
          if (aux_load && !PE) {
            const uint32_t j = slot ? aux_n1++ : aux_n0++;
            if (j > 0) mbar_wait(&aux_empty[slot], (j - 1) & 1u);
            if (elect_one()) {
              const int row0 = ((grp * 2 + slot) * 2 + (int)rank) * 128;
              if RN_PDBG(p, 4) { if (rank == 0) mbar_arrive(&aux_full[slot]); }
              else {
                if (rank == 0) mbar_arrive_expect_tx(&aux_full[slot], 2 * kAuxBytes);
                tma_load_2d_pair_hint(s_aux + slot * kAuxBytes, &p.tmAux[aux_load - 1], aux_full_leader + slot * 8, 0, row0, pol_s);
              }
            }
            __syncwarp();
          }
          // a layer whose chunks all fit the ring is loaded ONCE per group and used by both tile slots
          if (slot == 1 && k_chunks <= kNBStages) continue;
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(&empty_b[s], ph ^ 1);
            RN_TL(tl, 3000 + l * 10 + kc);                 // stage free, load issued
            if (elect_one()) {
              if RN_PDBG(p, 1) { if (rank == 0) mbar_arrive(&full_b[s]); }
              else {
                if (rank == 0) mbar_arrive_expect_tx(&full_b[s], (uint32_t)n * 128u);
                tma_load_2d_pair_hint(s_b + s * kBStageBytes, &p.tmB[l], full_b_leader + s * 8, kc * 64, (int)rank * (n >> 1), pol_w);
              }
            }
            __syncwarp();
            if (++s == kNBStages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA; whole warp converged, one elected lane issues) ----------------
    if (rank == 0) {
      constexpr uint32_t kHi = desc_hi_sw128(1024);
      const uint32_t act_lo0 = desc_lo_sw128(smem_u32(s_act), 16);
      const uint32_t aux_lo0 = desc_lo_sw128(smem_u32(s_aux), 16);
      const uint32_t b_lo0 = desc_lo_sw128(smem_u32(s_b), 16);
      int s = 0, s_mark = 0; uint32_t ph = 0, ph_mark = 0;
      uint32_t it0 = 0, it1 = 0, aux_n0 = 0, aux_n1 = 0;
      RN_TL_DECL(tl, 0, lane == 0);
      for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
        const int tiles_here = min(2, p.n_ptiles - grp * 2);
        for (int l = 0; l < p.n_layers; ++l) {
          const PairLayer& L = p.L[l];
          const uint32_t idesc = make_idesc_bf16(256, L.n, 0, 0);
          const int k_chunks = L.k_chunks;
          for (int slot = 0; slot < tiles_here; ++slot) {
            const uint32_t i = slot ? it1++ : it0++;
            RN_TL(tl, 100 + l * 10 + slot);                // item reached
            if (i > 0) mbar_wait(&act_ready[slot], (i - 1) & 1u);             // input tile written, accumulator drained
            RN_TL(tl, 300 + l * 10 + slot);                // input ready
            if (L.aux_load) {
              const uint32_t j = slot ? aux_n1++ : aux_n0++;
              mbar_wait(&aux_full[slot], j & 1u);
            }
            const uint32_t d_tmem = tmem_base + slot * 256;
            // weight stages of a layer that fits the ring are shared by the two slots: slot 0 consumes them without
            // releasing, slot 1 rewinds to the same stages and releases them
            const bool shared = (k_chunks <= kNBStages) && (tiles_here == 2);
            if (shared && slot == 1) { s = s_mark; ph = ph_mark; }
            s_mark = s; ph_mark = ph;
            for (int kc = 0; kc < k_chunks; ++kc) {
              if (!(shared && slot == 1)) mbar_wait(&full_b[s], ph);
              RN_TL(tl, 1000 + l * 100 + slot * 10 + kc);  // weight chunk present
              tcgen05_fence_after();
              if (elect_one()) {
                const int src = L.a_src[kc];
                const uint32_t al = (src == 4) ? aux_lo0 + slot * (kAuxBytes >> 4)
                                               : act_lo0 + slot * (kActBytes >> 4) + src * (kChunkBytes >> 4);
                const uint32_t bl = b_lo0 + s * (kBStageBytes >> 4);
                umma_bf16_pair(d_tmem, pack64(al, kHi), pack64(bl, kHi), idesc, kc != 0);
#pragma unroll
                for (int k = 1; k < 4; ++k) umma_bf16_pair(d_tmem, pack64(al + 2 * k, kHi), pack64(bl + 2 * k, kHi), idesc, 1u);
                if (!(shared && slot == 0)) umma_commit_pair(&empty_b[s]);
                if (kc == k_chunks - 1) {
                  umma_commit_pair(&tmem_full[slot]);
                  if (L.aux_release) umma_commit_pair(&aux_empty[slot]);
                }
              }
              __syncwarp();
              if (++s == kNBStages) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp < 18) {
    // ---------------- epilogue warps ----------------
    const int q = warp & 3;                       // TMEM lane quarter
    const int cq = (warp - 2) >> 2;               // column quarter of the layer output
    const int row = q * 32 + lane;
    const uint32_t act_ready_leader = mapa_u32(smem_u32(act_ready), 0);
    uint32_t it0 = 0, it1 = 0;
    RN_TL_DECL(tl, (warp == 2 ? 1 : 2), lane == 0 && (warp == 2 || warp == 17));
    const float2* consts2 = c_pair_consts2[p.const_slot];
    const float* consts1 = reinterpret_cast<const float*>(consts2);
    const int bt = (int)threadIdx.x - 64;             // 0..511 over the epilogue threads
    float pref0 = 0.f, pref1 = 0.f;                   // PE: next view-layer bias value of this thread, per tile slot
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      const int tiles_here = min(2, p.n_ptiles - grp * 2);
      for (int l = 0; l < p.n_layers; ++l) {
        const PairLayer L = p.L[l];
        const bool ray_bias = PE && (l == p.n_layers - 1);      // view layer: per-ray direction term instead of the bias
        if (!ray_bias) {
          // stage this layer's bias: every epilogue warp has finished the previous layer (first barrier), 256 threads copy
          // one value each out of the constant slot, and the copies are visible to all (second barrier)
          named_bar_sync(5, 512);
          if (bt < 256) s_bias[bt] = consts1[L.bias_off + bt];
          named_bar_sync(5, 512);
        }
        for (int slot = 0; slot < tiles_here; ++slot) {
          const uint32_t i = slot ? it1++ : it0++;
          const int64_t row0 = ((int64_t)(grp * 2 + slot) * 2 + rank) * 128;
          const int64_t gr = row0 + row;
          const bool row_ok = gr < p.m_rows;
          const float* sb = s_bias;
          if (PE && l >= p.n_layers - 2) {
            // the tile's rows belong to at most two rays (host-checked: group == 64 or >= 128): A = ray of its first row.
            // 32-bit arithmetic, last two layers only: a 64-bit division per item in every layer cost 3 % of the kernel.
            const uint32_t grp_pts = (uint32_t)p.group;
            const uint32_t rayA = (uint32_t)row0 / grp_pts;
            if (!ray_bias) {
              if (bt < 256) {
                // one layer ahead (thousands of cycles): fetch this thread's element of [dirvec[rayA] | dirvec[rayA + 1]]
                const uint32_t n_rays = (uint32_t)p.m_rows / grp_pts;
                const uint32_t ray = min(rayA + (uint32_t)(bt >> 7), n_rays - 1u);
                const float v = __ldg(p.dirvec + (size_t)ray * 128 + (bt & 127));
                if (slot) pref1 = v; else pref0 = v;
              }
            } else {
              named_bar_sync(5, 512);                  // all warps are done with the previous item's bias
              if (bt < 256) s_bias[bt] = slot ? pref1 : pref0;
              named_bar_sync(5, 512);
              const uint32_t rowsA = (rayA + 1u) * grp_pts - (uint32_t)row0;
              if ((uint32_t)row >= rowsA) sb = s_bias + 128;
            }
          }
          RN_TL(tl, 100 + l * 10 + slot);
          mbar_wait(&tmem_full[slot], i & 1u);
          RN_TL(tl, 300 + l * 10 + slot);                  // accumulator complete
          if (TRAIN && i > 0) mbar_wait(&store_done[slot], (i - 1) & 1u);   // the previous activation has been stored
          RN_TL(tl, 500 + l * 10 + slot);
          tcgen05_fence_after();
          uint8_t* s_tile = s_act + slot * kActBytes;
          const uint32_t t_addr = tmem_base + slot * 256 + ((uint32_t)(q * 32) << 16) + cq * (L.n >> 2);
          float h0 = 0.f, h1 = 0.f, h2 = 0.f;
          uint32_t mb[2] = {0u, 0u};
          if RN_PDBG(p, 16) { }
          else if RN_PDBG(p, 8) {
#pragma unroll 1
            for (int g = 0; g < (L.n >> 7); ++g) { uint32_t v[32]; tmem_ld_x32(t_addr + g * 32, v); tmem_ld_wait(); if (v[0] == 0x7fc12345u) mb[0] ^= v[1]; }
          }
          else {
            if (L.n == 256) {
              if (L.heads == 1) pair_epilogue<2, 1, true, TRAIN>(t_addr, s_tile, row, cq * 2, s_bias, consts2, L.head_w_off >> 1, 256, mb, h0, h1, h2);
              else if (L.relu) pair_epilogue<2, 0, true, TRAIN>(t_addr, s_tile, row, cq * 2, s_bias, consts2, L.head_w_off >> 1, 256, mb, h0, h1, h2);
              else pair_epilogue<2, 0, false, false>(t_addr, s_tile, row, cq * 2, s_bias, consts2, L.head_w_off >> 1, 256, mb, h0, h1, h2);
            } else {
              pair_epilogue<1, 3, true, false>(t_addr, s_tile, row, cq, sb, consts2, L.head_w_off >> 1, 128, mb, h0, h1, h2);
            }
          }
          RN_TL(tl, 600 + l * 10 + slot);                  // math done
          tcgen05_fence_before();
          if (!RN_PDBG(p, 128)) fence_proxy_async_smem();
          RN_TL(tl, 650 + l * 10 + slot);                  // fences done
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(act_ready_leader + slot * 8);
            if (TRAIN) mbar_arrive(&staged[slot]);
          }
          RN_TL(tl, 700 + l * 10 + slot);                  // tile written, arrived
          if (TRAIN && L.mask_out && row_ok)
            *reinterpret_cast<uint2*>(L.mask_out + gr * 8 + cq * 2) = make_uint2(mb[0], mb[1]);
          if (L.heads > 0) {
            // the four warps of a row quarter each hold a quarter of the head dot products: summed through shared memory
            // in a fixed order (column quarter 3, 2, 1, 0) -> deterministic
            const int nh = L.heads;
#pragma unroll 1
            for (int step = 3; step >= 1; --step) {
              if (cq == step) {
                s_hx[row] = (step == 3 ? 0.f : s_hx[row]) + h0;
                if (nh == 3) {
                  s_hx[128 + row] = (step == 3 ? 0.f : s_hx[128 + row]) + h1;
                  s_hx[256 + row] = (step == 3 ? 0.f : s_hx[256 + row]) + h2;
                }
              }
              named_bar_sync(1 + q, 128);
            }
            if (cq == 0 && row_ok) {
              const float* hb = consts1 + L.head_b_off;
              float* o = p.raw + gr * 4 + L.head_col;
              o[0] = h0 + s_hx[row] + hb[0];
              if (nh == 3) { o[1] = h1 + s_hx[128 + row] + hb[1]; o[2] = h2 + s_hx[256 + row] + hb[2]; }
            }
            named_bar_sync(1 + q, 128);
          }
        }
      }
    }
  } else if (warp == 18 && !PE) {
    // ---------------- store warp (training): activation tiles -> global for the backward pass ----------------
    if (TRAIN && lane == 0) {
      const uint64_t pol_s = l2_policy(p.l2_hints ? 1 : 0);
      uint32_t it0 = 0, it1 = 0;
      for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
        const int tiles_here = min(2, p.n_ptiles - grp * 2);
        for (int l = 0; l < p.n_layers; ++l) {
          const int chunks = (p.L[l].store && !RN_PDBG(p, 2)) ? (p.L[l].n >> 6) : 0;
          for (int slot = 0; slot < tiles_here; ++slot) {
            const uint32_t i = slot ? it1++ : it0++;
            const int row0 = ((grp * 2 + slot) * 2 + (int)rank) * 128;
            mbar_wait(&staged[slot], i & 1u);
            for (int c = 0; c < chunks; ++c)
              tma_store_2d_hint(&p.tmD[l], s_act + slot * kActBytes + c * kChunkBytes, c * 64, row0, pol_s);
            tma_store_commit();
            tma_store_wait_read0();
            mbar_arrive(&store_done[slot]);
          }
        }
      }
      tma_store_wait_all0();
    }
  } else if (PE && warp >= 18) {
    // ---------------- encoder warps (inference): x_enc of this CTA's 128 rows, straight into the side buffer ----------------
    // One warp per tile slot (warp 18, the store warp of the training kernel, <-> slot 0; warp 19 <-> slot 1; 640 threads
    // keep the 96-register budget -- at 672 ptxas sizes for 768 and squeezes the epilogue into 80): a tile costs one warp ~11,000 cycles and both
    // slots' buffers come free within one layer of each other, so a single warp encoding them back to back delivered the
    // second one late (first build: render 1.5 % SLOWER than with the separate encode kernel, profiles/r02_ab_log.md).
    // Same barrier protocol as the TMA-loaded side chunk: wait until layer 5's MMAs of the previous tile in this slot have
    // read the buffer (aux_empty), write the tile in the SWIZZLE_128B K-major layout, make it visible to the tensor core
    // (fence.proxy.async), arrive on the leader's aux_full (count 2: one per CTA).  The buffer is free from layer 5 on
    // (the direction term needs no K chunk any more), so the next group's tile is encoded under layers 6-9.
    const uint32_t aux_full_leader = mapa_u32(smem_u32(aux_full), 0);
    const int slot = warp - 18;
    uint32_t j = 0;
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      if (slot >= min(2, p.n_ptiles - grp * 2)) continue;
      if (j > 0) mbar_wait(&aux_empty[slot], (j - 1) & 1u);
      ++j;
      uint8_t* tile = s_aux + slot * kAuxBytes;
      const int64_t row0 = ((int64_t)(grp * 2 + slot) * 2 + rank) * 128;
      // Rolled (one frequency per iteration, 2-byte stores): ~340 SASS instructions instead of ~1,500 of straight-line
      // code next to the epilogue loop that paces the kernel; measured neutral against the unrolled version
      // (profiles/r02_ab_log.md, block 8), kept for its size.
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int r = i * 32 + lane;                 // lanes of a quarter warp hit eight different 16-byte columns
        const int64_t gr = row0 + r;
        float x[3] = {0.f, 0.f, 0.f};
        if (gr < p.m_rows) { x[0] = __ldg(p.pts + gr * 3); x[1] = __ldg(p.pts + gr * 3 + 1); x[2] = __ldg(p.pts + gr * 3 + 2); }
        uint8_t* rowp = tile + r * 128;
        const int sw = r & 7;
        auto put = [&](int f, float v) {             // feature f of this row -> its swizzled 2-byte slot
          *reinterpret_cast<__nv_bfloat16*>(rowp + (((f >> 3) ^ sw) << 4) + ((f & 7) << 1)) = __float2bfloat16_rn(v);
        };
        float t_hi[3], t_lo[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { put(a, x[a]); pe_turns(x[a], t_hi[a], t_lo[a]); }
        put(63, 0.f);
#pragma unroll 1
        for (int k = 0; k < kPosFreqs; ++k) {
          const float f = (float)(1 << k);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            float sn, cs;
            pe_sincos_turns(t_hi[a], t_lo[a], f, sn, cs);
            put(3 + 6 * k + a, sn);
            put(6 + 6 * k + a, cs);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(aux_full_leader + slot * 8);
    }
  }

  // ---------------- teardown: neither CTA may leave (or free TMEM) while its peer still works ----------------
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

// =====================================================================================================
// Data-gradient chain (backward of the trunk): dX_{l-1} = (dX_l W_l) .* (H_{l-1} > 0), nine layers in one launch.
//
// Same pair / two-slot / in-place structure as the forward chain.  Differences: the weights are read MN-major (the
// forward's [out][in] matrices serve as B[K = out][N = in] without a transposed copy); the epilogue applies the packed
// ReLU mask the forward stored (8 bytes per row and warp) instead of bias + ReLU; every layer's output is TMA-stored
// because the weight-gradient GEMMs read it afterwards; the first layer's input (dHC, 128 wide) is TMA-loaded into the
// activation buffer, and the sigma-gradient column block of dFS enters as a side chunk.  HBM traffic per point and layer:
// 512 B written (+ 8 x 4 B of mask read) instead of 512 B read + 512 B written by the per-layer kernel.
// =====================================================================================================
struct BwdLayer {
  int k_chunks;                // 64-wide K chunks of the A operand
  int8_t a_src[kPairMaxChunks];
  int last_ksteps;             // 16-wide k steps of the last chunk (1..4)
  int n_off;                   // first weight column (N offset inside the weight matrix)
  int aux_load, aux_col;       // side chunk: 0 / 1 + tensor map index, first column
  int act_load_chunks;         // >0: the layer input is TMA-loaded from tmIn into the activation buffer
  const uint32_t* mask_in;     // packed ReLU mask of the layer OUTPUT rows [M][8], or null
};
struct BwdParams {
  long long* timeline;          // RN_EXPERIMENTS only (scripts/pair_timeline.py bwd)
  CUtensorMap tmB[kPairMaxLayers], tmD[kPairMaxLayers], tmIn, tmAux;
  BwdLayer L[kPairMaxLayers];
  int n_layers, n_ptiles;
  int64_t m_rows;
  int l2_hints;
  uint32_t* flags;              // [n_layers][n_blocks] "block published" words for wgrad_stream.cu, or null
  int n_blocks;                 // 128-row blocks
  int stagger_ns;               // measurement only (rn_set_flag(11, us)): cluster c starts c / n_clusters of this window late
};

__device__ __forceinline__ void publish_block(uint32_t* f) {
  // The bulk stores of this block have COMPLETED (cp.async.bulk.wait_group, not .read: the writes are performed in L2,
  // the point of coherence every SM reads through), so a plain strong store of the flag is enough for a consumer that
  // acquires it.  A st.release.gpu here is a fence that also drains this thread's NEWER bulk stores -- 3-4 us per layer
  // and tile slot, which paced the whole chain at half its speed (profiles/r02_ab_log.md block 19).
  // The flag's value is the publication time (64 ns units, never 0): the consumer can report how long a block waited
  // for it (rn_debug_stream_lag) -- what decides whether it is still in L2.
  const uint32_t stamp = (uint32_t)(globaltimer_ns() >> 6) | 1u;
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(stamp) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
mlp_chain_pair_bwd_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_act = smem;
  uint8_t* s_aux = smem + 2 * kActBytes;
  uint8_t* s_b = s_aux + 2 * kAuxBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + kNBStages * kBStageBytes);
  uint64_t* full_b = bars;               // [4]  leader
  uint64_t* empty_b = bars + 4;          // [4]  each CTA (multicast commit)
  uint64_t* aux_full = bars + 8;         // [2]  leader
  uint64_t* aux_empty = bars + 10;       // [2]  each CTA
  uint64_t* act_ready = bars + 12;       // [2]  leader: 32 epilogue warps
  uint64_t* tmem_full = bars + 14;       // [2]  each CTA
  uint64_t* staged = bars + 16;          // [2]  epilogue -> store warp
  uint64_t* store_done = bars + 18;      // [2]  store warp -> epilogue and producer
  uint64_t* act_full = bars + 20;        // [2]  leader: TMA-loaded layer input landed in both CTAs
  uint64_t* act_free = bars + 22;        // [2]  store warp -> producer, ONCE PER TILE: the last layer's tile has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    for (int l = 0; l < p.n_layers; ++l) { prefetch_tmap(&p.tmB[l]); prefetch_tmap(&p.tmD[l]); }
    prefetch_tmap(&p.tmIn); prefetch_tmap(&p.tmAux);
    for (int i = 0; i < kNBStages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&aux_full[i], 1); mbar_init(&aux_empty[i], 1);
      mbar_init(&act_ready[i], 32); mbar_init(&tmem_full[i], 1);
      mbar_init(&staged[i], 16); mbar_init(&store_done[i], 1);
      mbar_init(&act_full[i], 1); mbar_init(&act_free[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const int n_groups = (p.n_ptiles + 1) >> 1;
  if (p.stagger_ns > 0) {
    const long long delay = (long long)p.stagger_ns * cluster_id / n_clusters;
    const unsigned long long t0 = globaltimer_ns();
    while ((long long)(globaltimer_ns() - t0) < delay) __nanosleep(500);
  }
  if (warp == 0) {
    // ---------------- TMA producer ----------------
    const uint32_t full_b_leader = mapa_u32(smem_u32(full_b), 0);
    const uint32_t aux_full_leader = mapa_u32(smem_u32(aux_full), 0);
    const uint32_t act_full_leader = mapa_u32(smem_u32(act_full), 0);
    int s = 0; uint32_t ph = 0;
    uint32_t aux_n0 = 0, aux_n1 = 0, in_n0 = 0, in_n1 = 0;
    const uint64_t pol_w = l2_policy(p.l2_hints ? 2 : 0), pol_s = l2_policy(p.l2_hints ? 1 : 0);
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      const int tiles_here = min(2, p.n_ptiles - grp * 2);
      for (int l = 0; l < p.n_layers; ++l) {
        const int k_chunks = p.L[l].k_chunks, aux_load = p.L[l].aux_load, act_chunks = p.L[l].act_load_chunks;
        const int n_off = p.L[l].n_off, aux_col = p.L[l].aux_col;
        for (int slot = 0; slot < tiles_here; ++slot) {
          const int row0 = ((grp * 2 + slot) * 2 + (int)rank) * 128;
          if (act_chunks) {
            // the buffer still feeds the TMA store of the previous tile's last layer.  (store_done completes once per
            // layer and this warp runs about a layer ahead of the stores: its parity would alias; act_free completes
            // once per tile.)
            const uint32_t j = slot ? in_n1++ : in_n0++;
            if (j > 0) mbar_wait(&act_free[slot], (j - 1) & 1u);
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(&act_full[slot], (uint32_t)(2 * act_chunks * kChunkBytes));
              for (int c = 0; c < act_chunks; ++c)
                tma_load_2d_pair_hint(s_act + slot * kActBytes + c * kChunkBytes, &p.tmIn, act_full_leader + slot * 8, c * 64, row0, pol_s);
            }
            __syncwarp();
          }
          if (aux_load) {
            const uint32_t j = slot ? aux_n1++ : aux_n0++;
            if (j > 0) mbar_wait(&aux_empty[slot], (j - 1) & 1u);
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(&aux_full[slot], 2 * kAuxBytes);
              tma_load_2d_pair_hint(s_aux + slot * kAuxBytes, &p.tmAux, aux_full_leader + slot * 8, aux_col, row0, pol_s);
            }
            __syncwarp();
          }
          if (slot == 1 && k_chunks <= kNBStages) continue;          // weight stages shared by the two slots
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(&empty_b[s], ph ^ 1);
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(&full_b[s], 2u * kBStageBytes);
              // this CTA's half of the 256 output columns: two [64 k rows][64 columns] boxes
              uint8_t* dst = s_b + s * kBStageBytes;
              const int c0 = n_off + (int)rank * 128;
              tma_load_2d_pair_hint(dst, &p.tmB[l], full_b_leader + s * 8, c0, kc * 64, pol_w);
              tma_load_2d_pair_hint(dst + 8192, &p.tmB[l], full_b_leader + s * 8, c0 + 64, kc * 64, pol_w);
            }
            __syncwarp();
            if (++s == kNBStages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA) ----------------
    if (rank == 0) {
      constexpr uint32_t kHi = desc_hi_sw128(1024);
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, 0, 1);
      const uint32_t act_lo0 = desc_lo_sw128(smem_u32(s_act), 16);
      const uint32_t aux_lo0 = desc_lo_sw128(smem_u32(s_aux), 16);
      const uint32_t b_lo0 = desc_lo_sw128(smem_u32(s_b), 8192);
      int s = 0, s_mark = 0; uint32_t ph = 0, ph_mark = 0;
      uint32_t it0 = 0, it1 = 0, aux_n0 = 0, aux_n1 = 0, in_n0 = 0, in_n1 = 0;
      RN_TL_DECL(tl, 0, lane == 0);
      for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
        const int tiles_here = min(2, p.n_ptiles - grp * 2);
        for (int l = 0; l < p.n_layers; ++l) {
          const BwdLayer& L = p.L[l];
          const int k_chunks = L.k_chunks;
          for (int slot = 0; slot < tiles_here; ++slot) {
            const uint32_t i = slot ? it1++ : it0++;
            RN_TL(tl, 100 + l * 10 + slot);
            if (i > 0) mbar_wait(&act_ready[slot], (i - 1) & 1u);             // accumulator drained (and input written)
            RN_TL(tl, 300 + l * 10 + slot);
            if (L.act_load_chunks) {
              const uint32_t j = slot ? in_n1++ : in_n0++;
              mbar_wait(&act_full[slot], j & 1u);
            }
            if (L.aux_load) {
              const uint32_t j = slot ? aux_n1++ : aux_n0++;
              mbar_wait(&aux_full[slot], j & 1u);
            }
            const uint32_t d_tmem = tmem_base + slot * 256;
            const bool shared = (k_chunks <= kNBStages) && (tiles_here == 2);
            if (shared && slot == 1) { s = s_mark; ph = ph_mark; }
            s_mark = s; ph_mark = ph;
            for (int kc = 0; kc < k_chunks; ++kc) {
              if (!(shared && slot == 1)) mbar_wait(&full_b[s], ph);
              RN_TL(tl, 1000 + l * 100 + slot * 10 + kc);
              tcgen05_fence_after();
              if (elect_one()) {
                const int src = L.a_src[kc];
                const uint32_t al = (src == 4) ? aux_lo0 + slot * (kAuxBytes >> 4)
                                               : act_lo0 + slot * (kActBytes >> 4) + src * (kChunkBytes >> 4);
                const uint32_t bl = b_lo0 + s * (kBStageBytes >> 4);
                const int ksteps = (kc == k_chunks - 1) ? L.last_ksteps : 4;
                umma_bf16_pair(d_tmem, pack64(al, kHi), pack64(bl, kHi), idesc, kc != 0);
#pragma unroll
                for (int k = 1; k < 4; ++k)
                  if (k < ksteps) umma_bf16_pair(d_tmem, pack64(al + 2 * k, kHi), pack64(bl + 128 * k, kHi), idesc, 1u);
                if (!(shared && slot == 0)) umma_commit_pair(&empty_b[s]);
                if (kc == k_chunks - 1) {
                  umma_commit_pair(&tmem_full[slot]);
                  if (L.aux_load) umma_commit_pair(&aux_empty[slot]);
                }
              }
              __syncwarp();
              if (++s == kNBStages) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp < 18) {
    // ---------------- epilogue warps: accumulator .* ReLU mask -> bf16 -> in-place tile ----------------
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t act_ready_leader = mapa_u32(smem_u32(act_ready), 0);
    uint32_t it0 = 0, it1 = 0;
    RN_TL_DECL(tl, (warp == 2 ? 1 : 2), lane == 0 && (warp == 2 || warp == 17));
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      const int tiles_here = min(2, p.n_ptiles - grp * 2);
      for (int l = 0; l < p.n_layers; ++l) {
        const uint32_t* mask_in = p.L[l].mask_in;
        for (int slot = 0; slot < tiles_here; ++slot) {
          const uint32_t i = slot ? it1++ : it0++;
          const int64_t gr = ((int64_t)(grp * 2 + slot) * 2 + rank) * 128 + row;
          uint2 mw = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
          if (mask_in && gr < p.m_rows) mw = __ldg(reinterpret_cast<const uint2*>(mask_in + gr * 8 + cq * 2));
          RN_TL(tl, 100 + l * 10 + slot);
          mbar_wait(&tmem_full[slot], i & 1u);
          RN_TL(tl, 300 + l * 10 + slot);
          if (i > 0) mbar_wait(&store_done[slot], (i - 1) & 1u);
          RN_TL(tl, 500 + l * 10 + slot);
          tcgen05_fence_after();
          uint8_t* s_tile = s_act + slot * kActBytes;
          const uint32_t t_addr = tmem_base + slot * 256 + ((uint32_t)(q * 32) << 16) + cq * 64;
          uint32_t v[2][32];
          tmem_ld_x32(t_addr, v[0]);
          tmem_ld_x32(t_addr + 32, v[1]);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int G = cq * 2 + g;
            const uint32_t word = g ? mw.y : mw.x;
            uint8_t* box = s_tile + (G >> 1) * kChunkBytes + row * 128;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int lchunk = (G & 1) * 4 + cc;
              uint4* dst = reinterpret_cast<uint4*>(box + ((lchunk ^ (row & 7)) << 4));
              uint32_t packed[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int pi = cc * 4 + e;                // column pair: flags at bits pi and 16 + pi of the word
                __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(v[g][2 * pi]), __uint_as_float(v[g][2 * pi + 1]));
                const uint32_t keep = ((word >> pi) & 0x00010001u) * 0xFFFFu;
                packed[e] = *reinterpret_cast<uint32_t*>(&pk) & keep;
              }
              *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
          }
          RN_TL(tl, 600 + l * 10 + slot);
          tcgen05_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(act_ready_leader + slot * 8);
            mbar_arrive(&staged[slot]);
          }
          RN_TL(tl, 700 + l * 10 + slot);
        }
      }
    }
  } else if (warp == 18) {
    // ---------------- store warp: every layer's data gradient goes to global for the weight-gradient GEMMs ----------------
    if (lane == 0) {
      // with a consumer beside this launch the gradients are re-read out of L2 within microseconds: keep them (normal
      // policy) instead of marking them evict_first
      const uint64_t pol_s = l2_policy(p.flags ? ((p.l2_hints & 2) ? 2 : 0) : ((p.l2_hints & 1) ? 1 : 0));
      uint32_t it0 = 0, it1 = 0;
      // flags of the last kPubLag store groups: a group is published kPubLag groups late, when cp.async.bulk.wait_group
      // says its writes have completed without this warp ever waiting for the most recent stores (a wait for the
      // previous group alone, ~3-4 us of write latency per layer and slot, paced the whole chain)
      constexpr int kPubLag = 4;
      uint32_t* pending[kPubLag] = {nullptr, nullptr, nullptr, nullptr};
      int pend_i = 0;
      RN_TL_DECL(tl, 3, true);
      for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
        const int tiles_here = min(2, p.n_ptiles - grp * 2);
        for (int l = 0; l < p.n_layers; ++l) {
          for (int slot = 0; slot < tiles_here; ++slot) {
            const uint32_t i = slot ? it1++ : it0++;
            const int row0 = ((grp * 2 + slot) * 2 + (int)rank) * 128;
            mbar_wait(&staged[slot], i & 1u);
            RN_TL(tl, 3300 + l * 10 + slot);
            for (int c = 0; c < 4; ++c)
              tma_store_2d_hint(&p.tmD[l], s_act + slot * kActBytes + c * kChunkBytes, c * 64, row0, pol_s);
            tma_store_commit();
            tma_store_wait_read0();
            RN_TL(tl, 3700 + l * 10 + slot);
            mbar_arrive(&store_done[slot]);
            if (l == p.n_layers - 1) mbar_arrive(&act_free[slot]);
            if (p.flags) {
              tma_store_wait_all4();                      // every group but the last kPubLag has completed its writes
              if (pending[pend_i]) publish_block(pending[pend_i]);
              pending[pend_i] = (row0 < p.m_rows) ? p.flags + (size_t)l * p.n_blocks + (row0 >> 7) : nullptr;
              pend_i = (pend_i + 1) & (kPubLag - 1);
            }
          }
        }
      }
      tma_store_wait_all0();
      if (p.flags)
        for (int i = 0; i < kPubLag; ++i)
          if (pending[i]) publish_block(pending[i]);
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

// Host launcher of the data-gradient chain.  Layer l: D_l[M,256] = (A_l[M,K_l] B_l[K_l, n_off : n_off+256]) .* mask_l with
// A_0 = `in` (in_cols wide, TMA-loaded), A_l = D_{l-1} (+ the side chunk `aux` at column aux_col as LAST K chunk when
// layers[l].aux_kind == 2).
int mlp_chain_pair_backward(const BwdLayerHost* layers, int n_layers, int64_t M, const void* in, int64_t ld_in, int in_cols,
                            const void* aux, int64_t ld_aux, int aux_cols, cudaStream_t st, uint32_t* flags, int n_sms) {
  int rc = check_arch();
  if (rc != RN_OK) return rc;
  RN_REQUIRE(layers && n_layers >= 1 && n_layers <= kPairMaxLayers && M > 0 && in && (in_cols == 128 || in_cols == 256));
  static thread_local BwdParams p;
  double flops = 0.0;
  if ((rc = make_tmap(&p.tmIn, in, in_cols, M, ld_in, 128)) != RN_OK) return rc;
  if ((rc = make_tmap(&p.tmAux, aux ? aux : in, aux ? aux_cols : in_cols, M, aux ? ld_aux : ld_in, 128)) != RN_OK) return rc;
  for (int l = 0; l < n_layers; ++l) {
    const BwdLayerHost& h = layers[l];
    RN_REQUIRE(h.B && h.D && h.k > 0 && h.k % 16 == 0 && h.k <= 320);
    // weights [k rows][ldb columns], MN-major B operand: boxes of [64 k rows][64 columns]
    if ((rc = make_tmap(&p.tmB[l], h.B, h.b_cols, h.k, h.ldb, 64)) != RN_OK) return rc;
    if ((rc = make_tmap(&p.tmD[l], h.D, 256, M, h.ldd, 128)) != RN_OK) return rc;
    BwdLayer& L = p.L[l];
    L = BwdLayer{};
    L.k_chunks = (h.k + 63) / 64;
    L.last_ksteps = ((h.k - (L.k_chunks - 1) * 64) + 15) / 16;
    int c = 0;
    for (int a = 0; c < L.k_chunks - (h.aux_kind == 2 ? 1 : 0); ++a) L.a_src[c++] = (int8_t)a;
    if (h.aux_kind == 2) { RN_REQUIRE(aux != nullptr); L.a_src[c++] = 4; L.aux_load = 1; L.aux_col = h.aux_col; }
    L.n_off = h.n_off;
    L.act_load_chunks = (l == 0) ? in_cols / 64 : 0;
    L.mask_in = h.mask_in;
    RN_REQUIRE(h.n_off + 256 <= h.b_cols);
    flops += 2.0 * (double)M * 256 * h.k;
  }
  p.n_layers = n_layers;
  p.n_ptiles = (int)ceil_div(M, 256);
  p.m_rows = M;
  p.l2_hints = (g_l2_hints & 1) | ((flags && !(g_l2_hints & 8)) ? 2 : 0);      // bit 1 here: evict_last stores, a consumer follows
  p.flags = flags;
  p.n_blocks = (int)ceil_div(M, 128);
  p.stagger_ns = flags ? g_ws_stagger_us * 1000 : 0;
#ifdef RN_EXPERIMENTS
  p.timeline = g_pair_timeline_bwd;
#else
  p.timeline = nullptr;
#endif
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    RN_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_pair_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
  const int n_groups = (p.n_ptiles + 1) / 2;
  int sms = (n_sms > 0 && n_sms < num_sms()) ? n_sms : num_sms();
  if (g_sm_limit_dgrad > 0 && g_sm_limit_dgrad < sms) sms = g_sm_limit_dgrad;
  const int max_clusters = sms / 2;
  const int grid = 2 * (n_groups < max_clusters ? n_groups : max_clusters);
  g_prof_next_flops = flops;
  int slot;
  prof_begin(1 /*MODE_NN*/, st, &slot);
  RN_CUDA_CHECK(launch_maybe_pdl(mlp_chain_pair_bwd_kernel, dim3(grid), dim3(kPairThreads), kPairSmem, st, p));
  prof_end(slot, st);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

// Constant-slot table, per device: slot <-> packed-weight buffer (key = address of its fp32 section).
struct ConstSlots {
  const void* key[kConstSlots] = {nullptr, nullptr, nullptr, nullptr};
  unsigned long long last_use[kConstSlots] = {0, 0, 0, 0};
  unsigned long long tick = 0;
};
static ConstSlots g_const_slots[kMaxDevices];
static std::mutex g_const_mu;

// Returns the slot of `consts`, uploading its current contents on `st` (stream-ordered before the launch that follows).
static int acquire_const_slot(const float* consts, cudaStream_t st, int* slot_out) {
  int dev = 0;
  RN_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_const_mu);
  ConstSlots& t = g_const_slots[dev & (kMaxDevices - 1)];
  int slot = -1, lru = 0;
  for (int i = 0; i < kConstSlots; ++i) {
    if (t.key[i] == consts) slot = i;
    if (t.last_use[i] < t.last_use[lru]) lru = i;
  }
  if (slot < 0) {
    slot = lru;
    if (t.key[slot] != nullptr) {
      // eviction: a kernel of the previous owner may still be reading the slot on another stream
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      RN_CUDA_CHECK(cudaStreamIsCapturing(st, &cs));
      if (cs != cudaStreamCaptureStatusNone) return RN_ERR_INVALID_ARG;   // > kConstSlots distinct networks inside one capture
      RN_CUDA_CHECK(cudaDeviceSynchronize());
    }
    t.key[slot] = consts;
  }
  t.last_use[slot] = ++t.tick;
  RN_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_pair_consts2, consts, layout::kF32Elems * sizeof(float),
                                        (size_t)slot * kConstSlotFloat2 * sizeof(float2), cudaMemcpyDeviceToDevice, st));
  *slot_out = slot;
  return RN_OK;
}

// Host launcher.  layers[l] uses the ChainLayerHost description of the per-layer chain (gemm.h); the pair kernel derives
// where each K chunk of A lives: activation chunks for the part produced by the previous layer, the side buffer for
// x_enc (first chunk of layers 0 and 5) and d_enc (last chunk of the view layer).
int mlp_chain_pair_forward(const ChainLayerHost* layers, int n_layers, int64_t M, const void* x_enc, int64_t ld_x,
                           const void* d_enc, int64_t ld_d, const float* consts, float* raw, bool training, cudaStream_t st,
                           const PairEncodeArgs* pe) {
  int rc = check_arch();
  if (rc != RN_OK) return rc;
  RN_REQUIRE(layers && n_layers >= 1 && n_layers <= kPairMaxLayers && M > 0 && consts && raw);
  RN_REQUIRE(pe ? (!training && pe->pts && pe->dirvec && pair_encode_supported(pe->group) && M % pe->group == 0 &&
                   M < (int64_t)0x7FFFFF00 && pe->group < (int64_t)0x7FFFFFFF) : (x_enc && d_enc));
  static thread_local PairParams p;     // ~4 KB of tensor maps: built on the host, passed by value as a kernel parameter
  double flops = 0.0;
  if (!pe) {
    if ((rc = make_tmap(&p.tmAux[0], x_enc, 64, M, ld_x, 128)) != RN_OK) return rc;
    if ((rc = make_tmap(&p.tmAux[1], d_enc, 64, M, ld_d, 128)) != RN_OK) return rc;
  } else {
    // never dereferenced in this mode (prefetch.tensormap only): any valid map
    if ((rc = make_tmap(&p.tmAux[0], layers[0].B, layers[0].k, layers[0].n, layers[0].ldb, layers[0].n / 2)) != RN_OK) return rc;
    p.tmAux[1] = p.tmAux[0];
  }
  p.pts = pe ? pe->pts : nullptr; p.dirvec = pe ? pe->dirvec : nullptr; p.group = pe ? pe->group : 1;
  for (int l = 0; l < n_layers; ++l) {
    const ChainLayerHost& h = layers[l];
    RN_REQUIRE((h.n == 256 || h.n == 128) && (h.k == 64 || h.k == 256 || h.k == 320));
    RN_REQUIRE(!(h.mask_out && h.n != 256) && !(h.n == 128 && h.heads != 3) && !(h.n == 256 && h.heads == 3));
    RN_REQUIRE(h.relu || (!h.mask_out && h.heads == 0));      // mask bits and fused heads assume a ReLU output
    RN_REQUIRE(h.bias_off % 2 == 0 && h.head_w_off % 2 == 0);
    if ((rc = make_tmap(&p.tmB[l], h.B, h.k, h.n, h.ldb, h.n / 2)) != RN_OK) return rc;
    if (training && (rc = make_tmap(&p.tmD[l], h.D, h.n, M, h.ldd, 128)) != RN_OK) return rc;
    PairLayer& L = p.L[l];
    L = PairLayer{};
    L.n = h.n; L.k_chunks = h.k / 64;
    // aux_kind: 0 none, 1 = x_enc is the FIRST chunk (layers 0 and 5), 2 = d_enc is the LAST chunk (view layer)
    int c = 0;
    if (h.aux_kind == 1) L.a_src[c++] = 4;
    for (int a = 0; c < L.k_chunks - (h.aux_kind == 2 ? 1 : 0); ++a) L.a_src[c++] = (int8_t)a;
    if (h.aux_kind == 2) L.a_src[c++] = 4;
    L.aux_load = h.aux_load; L.aux_release = h.aux_release;
    L.relu = h.relu; L.heads = h.heads; L.head_col = h.head_col;
    L.bias_off = h.bias_off; L.head_w_off = h.head_w_off; L.head_b_off = h.head_b_off;
    L.store = training ? 1 : 0;
    L.mask_out = training ? h.mask_out : nullptr;
    flops += 2.0 * (double)M * h.n * h.k;
  }
#ifdef RN_EXPERIMENTS
  p.dbg = g_chain_dbg;
  p.timeline = g_pair_timeline;
#else
  p.dbg = 0;
  p.timeline = nullptr;
#endif
  p.n_layers = n_layers;
  p.n_ptiles = (int)ceil_div(M, 256);
  p.m_rows = M; p.raw = raw; p.l2_hints = g_l2_hints & 1;
  if ((rc = acquire_const_slot(consts, st, &p.const_slot)) != RN_OK) return rc;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    RN_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    RN_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    RN_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
  }
  const int n_groups = (p.n_ptiles + 1) / 2;
  const int max_clusters = num_sms() / 2;
  const int grid = 2 * (n_groups < max_clusters ? n_groups : max_clusters);
  g_prof_next_flops = flops;
  int slot;
  prof_begin(0 /*MODE_NT*/, st, &slot);
  if (training) RN_CUDA_CHECK(launch_maybe_pdl(mlp_chain_pair_kernel<true, false>, dim3(grid), dim3(kPairThreads), kPairSmem, st, p));
  else if (pe) RN_CUDA_CHECK(launch_maybe_pdl(mlp_chain_pair_kernel<false, true>, dim3(grid), dim3(kPairThreadsPE), kPairSmem, st, p));
  else RN_CUDA_CHECK(launch_maybe_pdl(mlp_chain_pair_kernel<false, false>, dim3(grid), dim3(kPairThreads), kPairSmem, st, p));
  prof_end(slot, st);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

#ifdef RN_EXPERIMENTS
extern "C" int rn_pair_timeline(long long* device_buffer) { rn::g_pair_timeline = device_buffer; return 0; }
extern "C" int rn_pair_timeline_bwd(long long* device_buffer) { rn::g_pair_timeline_bwd = device_buffer; return 0; }
extern "C" int rn_pair_timeline_dims(int* roles, int* entries) { *roles = rn::kTlRoles; *entries = rn::kTlEntries; return 0; }
#endif
