// gemm.h -- host-side entry points of gemm_tcgen05.cu used by mlp.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rn {
struct TnInfo { int m_tiles, splits, N; float* scratch; };
int check_arch();
int gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int64_t M, int N, int K,
            const float* bias, int relu, cudaStream_t st, uint32_t* mask_out = nullptr, int n_heads = 0,
            const float* head_w = nullptr, const float* head_b = nullptr, float* head_out = nullptr, int head_col = 0);
int gemm_nn(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int64_t M, int N, int K,
            const uint32_t* mask_bits, cudaStream_t st);
size_t gemm_tn_scratch_bytes();
int gemm_tn_launch(const void* A, int64_t lda, int Mo, const void* B, int64_t ldb, int N, int64_t K, float* scratch,
                   size_t scratch_bytes, TnInfo* info, cudaStream_t st);
int gemm_tn_reduce(const TnInfo& info, int row0, int nrows, int col0, int ncols, float* dst, int64_t dst_ld,
                   float* colsum_dst, cudaStream_t st);
// Batched form: the backward pass gives every weight-gradient launch its own partial-tile region and reduces all of
// them (14 scatters per network) in ONE launch at the end instead of one small latency-bound launch each.
constexpr int kTnBatchMax = 16;
struct TnReduceDesc {
  const float* partial; float* dst; float* colsum_dst; int64_t dst_ld;
  int m_tiles, splits, BN, row0, nrows, col0, ncols, block0;
};
struct TnBatch { TnReduceDesc d[kTnBatchMax]; int n, total_blocks; };
int tn_batch_add(TnBatch* b, const TnInfo& info, int row0, int nrows, int col0, int ncols, float* dst, int64_t dst_ld,
                 float* colsum_dst);
int gemm_tn_reduce_batch(const TnBatch& b, cudaStream_t st);

// one layer of the chained forward (mlp_chain_pair_forward): D[M,n] = act(A[M,k] B[n,k]^T + bias), bf16 views
struct ChainLayerHost {
  const void* A; int64_t lda; const void* B; int64_t ldb; void* D; int64_t ldd;
  int k, n, relu, heads, head_col, bias_off, head_w_off, head_b_off, dep;
  uint32_t* mask_out;
  // pair kernel only (chain_pair.cu): which K chunk of A is a side chunk (0 none, 1 = first chunk is x_enc, 2 = last chunk
  // is d_enc), whether that side chunk is (re)loaded before this layer (1 = x_enc, 2 = d_enc) and released after it
  int aux_kind, aux_load, aux_release;
};
// PE-fused inference: the kernel encodes `pts` itself (no x_enc tensor), and the view layer takes dirvec[ray] as its bias
// (no d_enc K chunk: that layer's k is 256).  A 128-row tile may touch at most two rays.
struct PairEncodeArgs { const float* pts; const float* dirvec; int64_t group; };
inline bool pair_encode_supported(int64_t group) { return group == 64 || group >= 128; }
int mlp_chain_pair_forward(const ChainLayerHost* layers, int n_layers, int64_t M, const void* x_enc, int64_t ld_x,
                           const void* d_enc, int64_t ld_d, const float* consts, float* raw, bool training, cudaStream_t st,
                           const PairEncodeArgs* pe = nullptr);
// one layer of the data-gradient chain (chain_pair.cu): D[M,256] = (A[M,k] B[k, n_off : n_off+256]) .* mask
struct BwdLayerHost {
  const void* B; int64_t ldb; int b_cols;      // weight matrix [k rows][b_cols], leading dimension ldb
  int k, n_off;
  void* D; int64_t ldd;
  const uint32_t* mask_in;                     // packed ReLU mask of the rows of D, or null
  int aux_kind, aux_col;                       // 2: the last K chunk is the side chunk starting at column aux_col
};
// flags != null: the store warp publishes flags[l * ceil(M / 128) + block] = 1 once layer l's 128-row block is in global
// memory (for wgrad_stream.cu, which runs beside this launch); n_sms > 0 limits the launch to that many SMs.
int mlp_chain_pair_backward(const BwdLayerHost* layers, int n_layers, int64_t M, const void* in, int64_t ld_in, int in_cols,
                            const void* aux, int64_t ld_aux, int aux_cols, cudaStream_t st, uint32_t* flags = nullptr,
                            int n_sms = 0);

// wgrad_stream.cu: all weight / bias gradients of one backward pass in one persistent launch of CTA pairs that consumes
// the chain's output block by block.  Problem i: dW[Mo, N] = A[M pts, a_col0 : a_col0 + Mo]^T B[M pts, 0 : N] with Mo <= 256;
// a_cols / b_cols = columns the tensors really have (what lies beyond is read as zero); flag_row = the chain layer whose
// output is A (-1: A is complete before the launch).  `ctas` = SMs the launch may use; problem i's partial tiles go to
// scratch + i * region_floats and are described by infos[i] for tn_batch_add / gemm_tn_reduce_batch.
constexpr int kWsMaxProblems = 11;
struct WsHostProblem {
  const void* A; int64_t lda; int a_cols, a_col0, Mo;
  const void* B; int64_t ldb; int b_cols, N;
  int flag_row;
};
int wgrad_stream_plan(const WsHostProblem* probs, int n, int pairs, int* splits_out);
int wgrad_stream_launch(const WsHostProblem* probs, int n, int64_t M, const uint32_t* flags, int ctas, float* scratch,
                        size_t region_floats, TnInfo* infos, cudaStream_t st);

// inference: per-ray view-direction projection dirvec[R][128] (encode.cu: dir_bias_kernel)
int launch_dir_bias(const float* dirs, int64_t R, const void* packed, float* dirvec, cudaStream_t st);
// inference forward of one network (mlp.cu) and the per-view ray generator (raygen.cu), used by render.cu
size_t mlp_infer_workspace_bytes(int64_t M, int group);
int mlp_infer(const void* packed, const float* pts, const float* dirs, int64_t M, int group, void* ws, float* raw, cudaStream_t st);
int launch_view_rays(const float* pose, int W, float focal, float cx, float cy, int64_t ray0, int64_t B, float* ro, float* rd,
                     float* vd, cudaStream_t st);
int launch_encode(const float* pts, const float* dirs, int64_t M, int group, void* XC, int ldx, void* FD, int ldf, cudaStream_t st);
size_t heads_bwd_scratch_bytes(int64_t M);
int launch_heads_bwd(const float* g_raw, const void* HC, int64_t M, const float* f32sec, void* dHC, void* dFS, int ldfs,
                     float* scratch, float* gWrgb, float* gBrgb, cudaStream_t st, const void* H7 = nullptr,
                     float* gWsig = nullptr, float* gBsig = nullptr);
int launch_encode_bwd(const float* pts, const float* dirs, int64_t M, int group, const void* dXE0, const void* dXE5,
                      const void* dDE, float* g_pts, float* g_dirs, cudaStream_t st);
}  // namespace rn
