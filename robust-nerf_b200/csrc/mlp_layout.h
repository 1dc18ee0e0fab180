// mlp_layout.h -- packed (bf16, padded) weight cache and activation workspace layout of the
// reference's default NeRF (noisy_src/model.py:98-143): 8x256 trunk, skip after layer 4,
// sigma/feature heads, 283->128 view branch, 128->3 colour head.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace rn {
namespace layout {

// ---- packed weights: bf16 matrices (row-major [out][in_padded]) followed by an fp32 section ----
// W0p [256][64]   cols 0..62 = pts_linears.0.weight, col 63 = 0
// W1..W4,W6,W7 [256][256]
// W5p [256][320]  cols 0..62 = W5[:, :63] (x_enc part), col 63 = 0, cols 64..319 = W5[:, 63:] (h part)
// WFS [272][256]  rows 0..255 = feature_linear.weight, row 256 = sigma_linear.weight, rest 0
// WD  [128][320]  cols 0..255 = dir_linear.weight[:, :256], cols 256..282 = [:, 256:283], rest 0
constexpr size_t kW0 = 0;
constexpr size_t kW1 = kW0 + 256 * 64;           // element offsets (bf16)
constexpr size_t kW2 = kW1 + 256 * 256;
constexpr size_t kW3 = kW2 + 256 * 256;
constexpr size_t kW4 = kW3 + 256 * 256;
constexpr size_t kW5 = kW4 + 256 * 256;
constexpr size_t kW6 = kW5 + 256 * 320;
constexpr size_t kW7 = kW6 + 256 * 256;
constexpr size_t kWFS = kW7 + 256 * 256;
constexpr size_t kWD = kWFS + 272 * 256;
constexpr size_t kBf16Elems = kWD + 128 * 320;
constexpr size_t kBf16Bytes = kBf16Elems * 2;    // multiple of 1024
// fp32 section (float offsets from the start of the section)
constexpr size_t kB0 = 0;                        // 8 trunk biases, 256 each
constexpr size_t kBF = kB0 + 8 * 256;            // feature bias 256
constexpr size_t kBD = kBF + 256;                // dir bias 128
constexpr size_t kWSig = kBD + 128;              // sigma weight 256 (fp32 copy for the head kernel)
constexpr size_t kBSig = kWSig + 256;            // sigma bias (1, padded to 4)
constexpr size_t kWRgb = kBSig + 4;              // rgb weight [3][128]
constexpr size_t kBRgb = kWRgb + 384;            // rgb bias (3, padded to 4)
constexpr size_t kF32Elems = kBRgb + 4;
constexpr size_t kPackedBytes = kBf16Bytes + kF32Elems * 4;

constexpr size_t trunk_w(int l) {
  return l == 0 ? kW0 : l == 1 ? kW1 : l == 2 ? kW2 : l == 3 ? kW3 : l == 4 ? kW4 : l == 5 ? kW5 : l == 6 ? kW6 : kW7;
}

// ---- flat gradient buffer: state_dict order (model.py:119-143), RN_NUM_PARAMS floats ----
// pts_linears.0 (256x63 + 256), .1-.4 (256x256 + 256), .5 (256x319 + 256), .6-.7, sigma (1x256 + 1),
// feature (256x256 + 256), dir (128x283 + 128), rgb (3x128 + 3)
constexpr size_t kG_W0 = 0, kG_B0 = kG_W0 + 256 * 63;
constexpr size_t trunk_gw(int l) {
  // l in 0..7
  size_t off = 0;
  for (int i = 0; i < l; ++i) off += (size_t)256 * (i == 0 ? 63 : (i == 5 ? 319 : 256)) + 256;
  return off;
}
constexpr int trunk_in(int l) { return l == 0 ? 63 : (l == 5 ? 319 : 256); }
constexpr size_t trunk_gb(int l) { return trunk_gw(l) + (size_t)256 * trunk_in(l); }
constexpr size_t kG_WSig = trunk_gb(7) + 256;
constexpr size_t kG_BSig = kG_WSig + 256;
constexpr size_t kG_WF = kG_BSig + 1;
constexpr size_t kG_BF = kG_WF + 256 * 256;
constexpr size_t kG_WD = kG_BF + 256;
constexpr size_t kG_BD = kG_WD + 128 * 283;
constexpr size_t kG_WRgb = kG_BD + 128;
constexpr size_t kG_BRgb = kG_WRgb + 3 * 128;
constexpr size_t kG_Total = kG_BRgb + 3;
static_assert(kG_Total == 595844, "parameter count must match the reference (summary.json:46)");

// ---- packed ReLU masks: one 32-bit word per 32 consecutive columns of a row ----
// Column j = 2i + h (pair i = 0..15, half h) of the group sits at bit i (h = 0) or 16 + i (h = 1): the two flags of a
// packed bf16x2 word -- min.u16x2(word, 0x00010001) of the non-negative halves -- drop into place with one shift by i.
constexpr int relu_mask_bit(int j) { return (j & 1) ? 16 + (j >> 1) : (j >> 1); }

// ---- activation workspace (bf16 elements per point) ----
// training: XC[320] H0 H1 H2 H3 H5 H6 H7 (7 x 256) FD[320] HC[128] MB0..MB7 (packed masks) | backward: dHC[128] dFS[272] dA[256] dB[256]
//           dXE0[64] dXE5[64] dDE[64]   (dA/dB of the per-layer design became one buffer per layer, dH0..dH7[256]: the
//           chained data-gradient kernel produces all of them before the weight-gradient GEMMs read them)
constexpr int kTrainFwdElems = 320 + 7 * 256 + 320 + 128 + 8 * 16;  // 2688 (last term: 8 packed ReLU masks, 32 B/point each)
constexpr int kTrainBwdElems = 128 + 272 + 8 * 256 + 64 + 64 + 64;  // 2640: dHC dFS dH0..dH7 dXE0 dXE5 dDE
constexpr int kInferElems = 320 + 256 + 256 + 320 + 128;            // 1280

}  // namespace layout
}  // namespace rn
