// mlp.cu -- host-side orchestration of the NeRF MLP forward / backward
// (noisy_src/model.py:145-196 and its autograd backward) as a chain of tcgen05 GEMM launches over
// bf16 activation buffers that stay resident in HBM between forward and backward.
//
// Forward (per network, M points):
//   encode      pts -> XC[:, 0:64] (PE, bf16)   dirs -> FD[:, 256:320]
//   L0          XC[:, 0:64]  x W0p^T  -> H0            (K=64)
//   L1..L3      H(l-1)       x Wl^T   -> Hl
//   L4          H3           x W4^T   -> XC[:, 64:320] (the skip concat is a column offset, no copy)
//   L5          XC[:, 0:320] x W5p^T  -> H5            (K=320)
//   L6, L7      ...                   -> H7
//   feature     H7 x WFS[0:256]^T     -> FD[:, 0:256]  (no activation)
//   dir         FD[:, 0:320] x WD^T   -> HC            (N=128, K=320)
//   heads       sigma = H7 . w_sigma and rgb = HC . W_rgb^T are fp32 dot products fused into the
//               epilogues of L7 and of the view-branch GEMM -> raw[M,4]
// Backward: heads_bwd (dHC, dsigma, the rgb_linear and -- with the stream -- sigma_linear gradients), then ONE data-gradient
// chain launch (chain_pair.cu: dHC -> dF -> dH7 -> ... -> dH0 with the ReLU masks fused in) and BESIDE it, on its own SMs and
// a second stream, ONE weight-gradient launch for all ten GEMMs (wgrad_stream.cu), then one deterministic reduce of their
// partial tiles.  rn_set_flag(9, 0) / rn_set_flag(3, 0) fall back to one TN (split-K) / NN launch per layer
// (gemm_tcgen05.cu), which are also the cross-checks of the parity tests.
#include "common.cuh"
#include "gemm.h"
#include "mlp_layout.h"

namespace rn {
using namespace layout;
typedef __nv_bfloat16 bf16;

struct TrainWs {
  bf16 *XC, *H[8], *FD, *HC, *dHC, *dFS, *dH[8], *dXE0, *dXE5, *dDE;     // dH[l] = gradient w.r.t. the output of trunk layer l
  uint32_t* MB[8];          // packed ReLU masks of H0..H7, [M][8] words each (written by the forward epilogues)
  float* scratch;
  size_t scratch_bytes;
  float* heads_scratch;
  uint32_t* flags;          // [9][ceil(M / 128)] "block published" words: data-gradient chain -> weight-gradient stream
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
constexpr int kTnLaunches = 11;      // weight-gradient GEMMs per backward pass (10 launches, or 11 problems of the stream): each keeps its own partial-tile region

// inference: per-ray direction term of the view layer, [M / group][128] fp32 with group >= 64 (PE-fused chain)
static size_t flags_bytes(int64_t M) { return align_up((size_t)9 * (size_t)((M + 127) / 128) * sizeof(uint32_t), 256); }
static size_t dirvec_bytes(int64_t M) { return align_up((size_t)(M / 64 + 1) * 128 * sizeof(float), 256); }

extern int g_chain_fwd, g_pe_fused;
extern int g_l2_hints, g_sm_limit_dgrad, g_sm_limit_wgrad, g_prof_suppress;
extern double g_prof_next_flops;
void prof_begin(int mode, cudaStream_t st, int* slot);
void prof_end(int slot, cudaStream_t st);
static bool infer_fused(int64_t M, int group) {
  return g_chain_fwd == 2 && g_pe_fused && pair_encode_supported(group) && M < (int64_t)0x7FFFFF00;
}

// group > 0: the caller knows the direction grouping, and when the PE-fused chain will take it the workspace is just
// the per-ray direction term (8 B per point instead of 2,560)
static size_t workspace_bytes(int64_t M, int training, int group = 0) {
  if (!training && group > 0 && infer_fused(M, group)) return dirvec_bytes(M) + 256;
  const size_t elems = training ? (size_t)(kTrainFwdElems + kTrainBwdElems) : (size_t)kInferElems;
  size_t b = align_up((size_t)M * elems * 2, 256);
  if (!training) b += dirvec_bytes(M);
  if (training)
    b += kTnLaunches * align_up(gemm_tn_scratch_bytes(), 256) + align_up(heads_bwd_scratch_bytes(M), 256) + flags_bytes(M);
  return b + 256;
}

static void carve_train(void* ws, int64_t M, TrainWs* w) {
  bf16* p = reinterpret_cast<bf16*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));
  auto take = [&](int width) { bf16* r = p; p += (size_t)M * width; return r; };
  w->XC = take(320);
  w->H[0] = take(256); w->H[1] = take(256); w->H[2] = take(256); w->H[3] = take(256);
  w->H[4] = w->XC + 64;                           // H4 lives in XC[:, 64:320] (ld 320)
  w->H[5] = take(256); w->H[6] = take(256); w->H[7] = take(256);
  w->FD = take(320); w->HC = take(128);
  for (int i = 0; i < 8; ++i) w->MB[i] = reinterpret_cast<uint32_t*>(take(16));
  w->dHC = take(128); w->dFS = take(272);
  for (int i = 0; i < 8; ++i) w->dH[i] = take(256);
  w->dXE0 = take(64); w->dXE5 = take(64); w->dDE = take(64);
  uint8_t* q = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(p), 256));
  w->scratch = reinterpret_cast<float*>(q);
  w->scratch_bytes = gemm_tn_scratch_bytes();
  w->heads_scratch = reinterpret_cast<float*>(q + kTnLaunches * align_up(w->scratch_bytes, 256));
  w->flags = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(w->heads_scratch) + align_up(heads_bwd_scratch_bytes(M), 256));
}

// ---- the side stream the weight-gradient consumer runs on (one per device, created on first use) ----
struct SideStream { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static int side_stream(SideStream** out) {
  static SideStream table[kMaxDevices];
  int dev = 0;
  RN_CUDA_CHECK(cudaGetDevice(&dev));
  SideStream& s = table[dev & (kMaxDevices - 1)];
  if (!s.st) {
    RN_CUDA_CHECK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    RN_CUDA_CHECK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    RN_CUDA_CHECK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return RN_OK;
}

static inline int ld_of(int l) { return l == 4 ? 320 : 256; }   // leading dimension of H[l]

// The ten weight-gradient GEMMs of one backward pass as problems of the stream (wgrad_stream.cu), in the order their
// partial-tile regions and reduce descriptors use.  Chain layer c writes: 0 -> dFS[:, 0:256], 1 -> dH7, ..., 8 -> dH0; dHC
// comes from heads_bwd (complete before the launch), which also sums the one-row sigma_linear gradient.
static int stream_problems(const TrainWs& w, WsHostProblem* P) {
  int np = 0;
  P[np++] = WsHostProblem{w.dHC, 128, 128, 0, 128, w.FD, 320, 320, 320, -1};           // dir_linear
  P[np++] = WsHostProblem{w.dFS, 272, 272, 0, 256, w.H[7], 256, 256, 256, 0};           // feature_linear
  for (int l = 7; l >= 1; --l)
    P[np++] = (l == 5) ? WsHostProblem{w.dH[5], 256, 256, 0, 256, w.XC, 320, 320, 320, 8 - 5}
                       : WsHostProblem{w.dH[l], 256, 256, 0, 256, w.H[l - 1], ld_of(l - 1), 256, 256, 8 - l};
  P[np++] = WsHostProblem{w.dH[0], 256, 256, 0, 256, w.XC, 320, 64, 64, 8};             // layer 0: x_enc only
  return np;
}

#define RN_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != RN_OK) return _rc; \
  } while (0)

#ifdef RN_EXPERIMENTS
int g_chain_dbg = 0;
#endif
// rn_set_flag(9, n): the weight gradients run BESIDE the data-gradient chain on n SMs (wgrad_stream.cu), taking each block
// of dH as the chain publishes it.  -1 (default) = 84 of 148 SMs, the measured optimum (42 pairs: 5 splits for dir_linear
// and layer 5, 4 for the other eight GEMMs; the chain keeps 64; 82 measures the same), off on a part with another SM
// count; 0 = off: one split-K launch per layer after the chain (more time per step, profiles/r02_ab_log.md blocks 19-26).
int g_wgrad_stream_sms = -1;
static int wgrad_stream_sms() {
  if (g_wgrad_stream_sms >= 0) return g_wgrad_stream_sms;
  return num_sms() == 148 ? 84 : 0;
}
// rn_set_flag(10, mask): measurement only.  bit 0 = record the two launches separately as well as the span; bit 1 = launch
// the stream AFTER the chain on the same stream (every flag already set: the consumer alone, operands from HBM)
int g_ws_debug = 0;
int g_ws_stagger_us = 0;      // rn_set_flag(11, us): measurement only, staggered start of the chain's clusters
int g_chain_bwd = 1;      // rn_set_flag(3, v): 1 = data gradients as one CTA-pair chain launch (chain_pair.cu), 0 = one launch per layer
// rn_set_flag(0, v): 0 = one launch per layer (gemm_tcgen05.cu: the building-block kernels, kept as the cross-check of the
// chain), 2 = CTA-pair chain with shared-memory-resident activations (chain_pair.cu; default).  (Round 1's third variant,
// a layer-chained launch whose activations round-tripped through L2, was measured slower than the pair chain on every
// shape and has been removed.)
int g_chain_fwd = 2;
// rn_set_flag(4, v): 1 = inference encodes the points inside the forward chain and hoists the view-direction term per ray
// (default), 0 = separate encode kernel + TMA-loaded x_enc / d_enc chunks (the training layout; kept for A/B and as the
// path for direction groups the fused kernel does not cover)
int g_pe_fused = 1;

static int mlp_forward(const void* packed, const float* pts, const float* dirs, int64_t M, int group, void* ws,
                       int training, float* raw, cudaStream_t st) {
  const bf16* W = reinterpret_cast<const bf16*>(packed);
  const float* F = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + kBf16Bytes);
  bf16 *XC, *H[8], *FD, *HC;
  uint32_t* MB[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (training) {
    TrainWs w;
    carve_train(ws, M, &w);
    XC = w.XC; FD = w.FD; HC = w.HC;
    for (int i = 0; i < 8; ++i) { H[i] = w.H[i]; MB[i] = w.MB[i]; }
  } else {
    bf16* p = reinterpret_cast<bf16*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));
    XC = p; p += (size_t)M * 320;
    bf16* HA = p; p += (size_t)M * 256;
    bf16* HB = p; p += (size_t)M * 256;
    FD = p; p += (size_t)M * 320;
    HC = p;
    H[0] = HA; H[1] = HB; H[2] = HA; H[3] = HB; H[4] = XC + 64; H[5] = HA; H[6] = HB; H[7] = HA;
  }
  const bool pe_fused = !training && infer_fused(M, group);
  float* dirvec = nullptr;
  if (pe_fused) {
    dirvec = reinterpret_cast<float*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));      // the only workspace this path touches
    RN_TRY(launch_dir_bias(dirs, M / group, packed, dirvec, st));
  } else {
    RN_TRY(launch_encode(pts, dirs, M, group, XC, 320, FD, 320, st));
  }
  if (g_chain_fwd) {
    ChainLayerHost L[10];
    for (int l = 0; l < 8; ++l) {
      const bool skip_in = (l == 5);
      L[l] = ChainLayerHost{l == 0 ? (const void*)XC : (skip_in ? (const void*)XC : (const void*)H[l - 1]),
                            (l == 0 || skip_in) ? 320 : ld_of(l - 1), W + trunk_w(l), l == 0 ? 64 : (skip_in ? 320 : 256),
                            H[l], ld_of(l), l == 0 ? 64 : (skip_in ? 320 : 256), 256, 1, l == 7 ? 1 : 0, 3,
                            (int)(kB0 + 256 * l), (int)kWSig, (int)kBSig, l == 0 ? 0 : 1, MB[l]};
    }
    L[8] = ChainLayerHost{H[7], 256, W + kWFS, 256, FD, 320, 256, 256, 0, 0, 0, (int)kBF, 0, 0, 1, nullptr};
    L[9] = ChainLayerHost{FD, 320, W + kWD, 320, HC, 128, 320, 128, 1, 3, 0, (int)kBD, (int)kWRgb, (int)kBRgb, 1, nullptr};
    if (g_chain_fwd == 2) {
      L[0].aux_kind = 1; L[0].aux_load = 1;
      L[5].aux_kind = 1; L[5].aux_release = 1;
      if (pe_fused) {
        L[9].k = 256;                                   // the d_enc K chunk became the per-ray bias dirvec
        const PairEncodeArgs pe{pts, dirvec, group};
        return mlp_chain_pair_forward(L, 10, M, nullptr, 0, nullptr, 0, F, raw, false, st, &pe);
      }
      L[9].aux_kind = 2; L[9].aux_load = 2; L[9].aux_release = 1;
      return mlp_chain_pair_forward(L, 10, M, XC, 320, FD + 256, 320, F, raw, training != 0, st);
    }
    return RN_ERR_INVALID_ARG;        // unreachable: g_chain_fwd is 0 or 2
  }
  RN_TRY(gemm_nt(XC, 320, W + kW0, 64, H[0], 256, M, 256, 64, F + kB0, 1, st, MB[0]));
  for (int l = 1; l < 8; ++l) {
    const bf16* in = (l == 5) ? XC : H[l - 1];
    const int K = (l == 5) ? 320 : 256;
    const int ldin = (l == 5) ? 320 : ld_of(l - 1);
    if (l == 7)   // sigma head fused into the epilogue: raw[:, 3] = H7 . w_sigma + b_sigma
      RN_TRY(gemm_nt(in, ldin, W + trunk_w(l), K, H[l], ld_of(l), M, 256, K, F + kB0 + 256 * l, 1, st, MB[l], 1, F + kWSig,
                     F + kBSig, raw, 3));
    else
      RN_TRY(gemm_nt(in, ldin, W + trunk_w(l), K, H[l], ld_of(l), M, 256, K, F + kB0 + 256 * l, 1, st, MB[l]));
  }
  RN_TRY(gemm_nt(H[7], 256, W + kWFS, 256, FD, 320, M, 256, 256, F + kBF, 0, st));
  // rgb head fused into the view-branch epilogue: raw[:, 0:3] = HC . W_rgb^T + b_rgb
  RN_TRY(gemm_nt(FD, 320, W + kWD, 320, HC, 128, M, 128, 320, F + kBD, 1, st, nullptr, 3, F + kWRgb, F + kBRgb, raw, 0));
  return RN_OK;
}

static int mlp_backward(const void* packed, const float* pts, const float* dirs, int64_t M, int group, void* ws,
                        const float* g_raw, float* G, float* g_pts, float* g_dirs, cudaStream_t st) {
  const bf16* W = reinterpret_cast<const bf16*>(packed);
  const float* F = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + kBf16Bytes);
  TrainWs w;
  carve_train(ws, M, &w);
  const bool need_in = (g_pts != nullptr) || (g_dirs != nullptr);
  TnInfo ti;
  TnBatch batch{};
  int tn_k = 0;
  const size_t region = align_up(w.scratch_bytes, 256) / sizeof(float);
  auto scratch_k = [&]() { return w.scratch + (size_t)(tn_k++) * region; };
  const int stream_sms = wgrad_stream_sms();
  const bool overlapped = g_chain_bwd && stream_sms >= 24 && stream_sms <= num_sms() - 24;
  // heads: dHC, dFS[:, 256:272], rgb_linear grads -- and, when the weight-gradient stream runs, the sigma_linear gradients
  // too (one row of dFS^T: a CTA pair per split in the stream, 256 B per point of extra reads here)
  RN_TRY(launch_heads_bwd(g_raw, w.HC, M, F, w.dHC, w.dFS, 272, w.heads_scratch, G + kG_WRgb, G + kG_BRgb, st,
                          overlapped ? w.H[7] : nullptr, G + kG_WSig, G + kG_BSig));
  if (g_chain_bwd) {
    // ---- all data gradients in one launch: dHC -> dF -> dH7 -> ... -> dH0 ----
    BwdLayerHost L[9];
    L[0] = BwdLayerHost{W + kWD, 320, 320, 128, 0, w.dFS, 272, nullptr, 0, 0};            // d feat = dHC x WD[:, 0:256]
    L[1] = BwdLayerHost{W + kWFS, 256, 256, 272, 0, w.dH[7], 256, w.MB[7], 2, 256};       // dH7 = [dF | dsigma] x WFS .* mask
    for (int l = 7; l >= 1; --l) {
      const bool skip = (l == 5);                                                         // dH4 = dH5 x W5[:, 63:] (packed cols 64..)
      L[9 - l] = BwdLayerHost{W + trunk_w(l), skip ? 320 : 256, skip ? 320 : 256, 256, skip ? 64 : 0, w.dH[l - 1], 256,
                              w.MB[l - 1], 0, 0};
    }
    if (overlapped) {
      // ---- ... with every weight gradient in a second launch beside it (wgrad_stream.cu) ----
      // chain layer c writes: 0 -> dFS[:, 0:256], 1 -> dH7, ..., 8 -> dH0; dHC comes from heads_bwd (complete already)
      WsHostProblem P[kWsMaxProblems];
      const int np = stream_problems(w, P);
      const int ws_sms = stream_sms & ~1;
      SideStream* side;
      RN_TRY(side_stream(&side));
      RN_CUDA_CHECK(cudaMemsetAsync(w.flags, 0, flags_bytes(M), st));
      // measurement hook: the two launches are timed as ONE span on the main stream (their own records would overlap)
      double span_flops = 0.0;
      for (int l = 0; l < 9; ++l) span_flops += 2.0 * (double)M * 256 * L[l].k;
      for (int i = 0; i < np; ++i) span_flops += 2.0 * (double)M * P[i].N * P[i].Mo;
      g_prof_next_flops = span_flops;
      int slot;
      prof_begin(3, st, &slot);
      // fork; once the side stream depends on the caller's stream it MUST join again, whatever fails in between
      RN_CUDA_CHECK(cudaEventRecord(side->fork, st));
      RN_CUDA_CHECK(cudaStreamWaitEvent(side->st, side->fork, 0));
      if (!(g_ws_debug & 1)) ++g_prof_suppress;
      int rc = mlp_chain_pair_backward(L, 9, M, w.dHC, 128, 128, w.dFS, 272, 272, st, w.flags, num_sms() - ws_sms);
      TnInfo infos[kWsMaxProblems];
      if (rc == RN_OK) rc = wgrad_stream_launch(P, np, M, w.flags, ws_sms, w.scratch, region, infos, (g_ws_debug & 2) ? st : side->st);
      if (!(g_ws_debug & 1)) --g_prof_suppress;
      // join even after a failed launch: a stream capture must not end with the side stream still forked
      cudaError_t e1 = cudaEventRecord(side->join, side->st);
      cudaError_t e2 = cudaStreamWaitEvent(st, side->join, 0);
      prof_end(slot, st);
      RN_TRY(rc);
      RN_CUDA_CHECK(e1);
      RN_CUDA_CHECK(e2);
      int k = 0;
      RN_TRY(tn_batch_add(&batch, infos[k++], 0, 128, 0, 283, G + kG_WD, 283, G + kG_BD));
      RN_TRY(tn_batch_add(&batch, infos[k++], 0, 256, 0, 256, G + kG_WF, 256, G + kG_BF));
      for (int l = 7; l >= 1; --l) {
        if (l == 5) {
          RN_TRY(tn_batch_add(&batch, infos[k], 0, 256, 0, 63, G + trunk_gw(5), 319, nullptr));
          RN_TRY(tn_batch_add(&batch, infos[k++], 0, 256, 64, 256, G + trunk_gw(5) + 63, 319, G + trunk_gb(5)));
        } else {
          RN_TRY(tn_batch_add(&batch, infos[k++], 0, 256, 0, 256, G + trunk_gw(l), 256, G + trunk_gb(l)));
        }
      }
      RN_TRY(tn_batch_add(&batch, infos[k++], 0, 256, 0, 63, G + trunk_gw(0), 63, G + trunk_gb(0)));
    } else {
      RN_TRY(mlp_chain_pair_backward(L, 9, M, w.dHC, 128, 128, w.dFS, 272, 272, st));
    }
  } else {
    RN_TRY(gemm_nn(w.dHC, 128, W + kWD, 320, w.dFS, 272, M, 256, 128, nullptr, st));
    RN_TRY(gemm_nn(w.dFS, 272, W + kWFS, 256, w.dH[7], 256, M, 256, 272, w.MB[7], st));
    for (int l = 7; l >= 1; --l)
      RN_TRY(gemm_nn(w.dH[l], 256, W + trunk_w(l) + (l == 5 ? 64 : 0), l == 5 ? 320 : 256, w.dH[l - 1], 256, M, 256, 256,
                     w.MB[l - 1], st));
  }
  if (!overlapped) {
    // ---- weight / bias gradients: one split-K GEMM per layer, reduced together below ----
    // dir_linear: one launch over the whole 320-wide input [feat(256) | d_enc(27) | 0]
    RN_TRY(gemm_tn_launch(w.dHC, 128, 128, w.FD, 320, 320, M, scratch_k(), w.scratch_bytes, &ti, st));
    RN_TRY(tn_batch_add(&batch, ti, 0, 128, 0, 283, G + kG_WD, 283, G + kG_BD));
    // feature_linear + sigma_linear (row 256 of dFS^T)
    RN_TRY(gemm_tn_launch(w.dFS, 272, 272, w.H[7], 256, 256, M, scratch_k(), w.scratch_bytes, &ti, st));
    RN_TRY(tn_batch_add(&batch, ti, 0, 256, 0, 256, G + kG_WF, 256, G + kG_BF));
    RN_TRY(tn_batch_add(&batch, ti, 256, 1, 0, 256, G + kG_WSig, 256, G + kG_BSig));
    for (int l = 7; l >= 1; --l) {
      if (l == 5) {
        // input = XC = [x_enc(64) | H4(256)]: one launch over the 320-wide input; the zero pad column 63 is dropped
        RN_TRY(gemm_tn_launch(w.dH[5], 256, 256, w.XC, 320, 320, M, scratch_k(), w.scratch_bytes, &ti, st));
        RN_TRY(tn_batch_add(&batch, ti, 0, 256, 0, 63, G + trunk_gw(5), 319, nullptr));
        RN_TRY(tn_batch_add(&batch, ti, 0, 256, 64, 256, G + trunk_gw(5) + 63, 319, G + trunk_gb(5)));
      } else {
        RN_TRY(gemm_tn_launch(w.dH[l], 256, 256, w.H[l - 1], ld_of(l - 1), 256, M, scratch_k(), w.scratch_bytes, &ti, st));
        RN_TRY(tn_batch_add(&batch, ti, 0, 256, 0, 256, G + trunk_gw(l), 256, G + trunk_gb(l)));
      }
    }
    // layer 0: input = x_enc
    RN_TRY(gemm_tn_launch(w.dH[0], 256, 256, w.XC, 320, 64, M, scratch_k(), w.scratch_bytes, &ti, st));
    RN_TRY(tn_batch_add(&batch, ti, 0, 256, 0, 63, G + trunk_gw(0), 63, G + trunk_gb(0)));
    RN_REQUIRE(tn_k <= kTnLaunches);
  }
  RN_TRY(gemm_tn_reduce_batch(batch, st));
  if (need_in) {
    // gradients w.r.t. the encodings (pose optimisation only): three 64-wide data-gradient GEMMs
    if (g_dirs) RN_TRY(gemm_nn(w.dHC, 128, W + kWD + 256, 320, w.dDE, 64, M, 64, 128, nullptr, st));
    RN_TRY(gemm_nn(w.dH[5], 256, W + kW5, 320, w.dXE5, 64, M, 64, 256, nullptr, st));
    RN_TRY(gemm_nn(w.dH[0], 256, W + kW0, 64, w.dXE0, 64, M, 64, 256, nullptr, st));
    RN_TRY(launch_encode_bwd(pts, dirs, M, group, w.dXE0, w.dXE5, w.dDE, g_pts, g_dirs, st));
  }
  return RN_OK;
}

// inference entry points for render.cu
size_t mlp_infer_workspace_bytes(int64_t M, int group) { return M > 0 ? workspace_bytes(M, 0, group) : 0; }
int mlp_infer(const void* packed, const float* pts, const float* dirs, int64_t M, int group, void* ws, float* raw, cudaStream_t st) {
  if (M == 0) return RN_OK;
  RN_TRY(check_arch());
  return mlp_forward(packed, pts, dirs, M, group, ws, 0, raw, st);
}

}  // namespace rn

using namespace rn;

extern "C" {

int rn_set_flag(int flag, int value) {
  if (flag == 0) { if (value != 0 && value != 2) return RN_ERR_INVALID_ARG; g_chain_fwd = value; return RN_OK; }
#ifdef RN_EXPERIMENTS
  if (flag == 1) { g_chain_dbg = value; return RN_OK; }
#endif
  if (flag == 3) { g_chain_bwd = value; return RN_OK; }
  if (flag == 4) { g_pe_fused = value; return RN_OK; }
  if (flag == 5) { g_pdl = value ? 1 : 0; return RN_OK; }
  if (flag == 6) { g_l2_hints = value & 11; return RN_OK; }
  if (flag == 7) { g_sm_limit_dgrad = value; return RN_OK; }
  if (flag == 8) { g_sm_limit_wgrad = value; return RN_OK; }
  if (flag == 9) { g_wgrad_stream_sms = value; return RN_OK; }
  if (flag == 10) { g_ws_debug = value; return RN_OK; }
  if (flag == 11) { g_ws_stagger_us = value; return RN_OK; }
  return RN_ERR_INVALID_ARG;
}

int rn_get_flag(int flag, int* value_host) {
  RN_REQUIRE(value_host);
  switch (flag) {
    case 0: *value_host = g_chain_fwd; return RN_OK;
    case 3: *value_host = g_chain_bwd; return RN_OK;
    case 4: *value_host = g_pe_fused; return RN_OK;
    case 5: *value_host = g_pdl; return RN_OK;
    case 6: *value_host = g_l2_hints; return RN_OK;
    case 7: *value_host = g_sm_limit_dgrad; return RN_OK;
    case 8: *value_host = g_sm_limit_wgrad; return RN_OK;
    case 9: *value_host = wgrad_stream_sms(); return RN_OK;
    case 10: *value_host = g_ws_debug; return RN_OK;
    case 11: *value_host = g_ws_stagger_us; return RN_OK;
    default: return RN_ERR_INVALID_ARG;
  }
}

int rn_debug_stream_plan(int sms, int* splits_out, int* n_problems_out) {
  // host logic only (no device work): how the weight-gradient stream would split its GEMMs over `sms` SMs
  RN_REQUIRE(splits_out && n_problems_out);
  TrainWs w{};
  WsHostProblem P[kWsMaxProblems];
  const int np = stream_problems(w, P);
  *n_problems_out = np;
  return wgrad_stream_plan(P, np, sms / 2, splits_out);
}

size_t rn_mlp_workspace_bytes(int64_t M, int training) { return M > 0 ? workspace_bytes(M, training) : 0; }
size_t rn_mlp_infer_workspace_bytes(int64_t M, int dir_group) { return mlp_infer_workspace_bytes(M, dir_group); }

int rn_mlp_fwd(const void* packed, const float* pts, const float* dirs, int64_t M, int dir_group, void* workspace,
               int training, float* raw_out, rn_stream_t stream) {
  if (M == 0) return RN_OK;
  RN_REQUIRE(packed && pts && dirs && workspace && raw_out && M >= 0 && dir_group >= 1 && M % dir_group == 0);
  RN_TRY(check_arch());
  return mlp_forward(packed, pts, dirs, M, dir_group, workspace, training, raw_out, (cudaStream_t)stream);
}

int rn_mlp_bwd(const void* packed, const float* pts, const float* dirs, int64_t M, int dir_group, void* workspace,
               const float* g_raw, float* grad_flat, float* g_pts, float* g_dirs, rn_stream_t stream) {
  RN_REQUIRE(packed && pts && dirs && workspace && g_raw && grad_flat && M > 0 && dir_group >= 1 && M % dir_group == 0);
  RN_TRY(check_arch());
  return mlp_backward(packed, pts, dirs, M, dir_group, workspace, g_raw, grad_flat, g_pts, g_dirs, (cudaStream_t)stream);
}

}  // extern "C"
