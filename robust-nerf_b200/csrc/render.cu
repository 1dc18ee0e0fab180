// render.cu -- evaluation rendering in ONE C-ABI call per tile range of a view.
//
// Replaces the chunk loop of noisy_src/rendering.py:287-323 around render_rays (rendering.py:119-240, is_train=False) and
// the per-image ray generation of noisy_src/train.py:122-160 / inference.py:76-105: rays are generated from the camera
// inside the call (no (H, W, 3) direction table, no per-view ray tensors), and the whole coarse -> resample -> fine
// pipeline of every tile is enqueued by this function -- nine kernel launches per tile, no Python between them, no host
// synchronisation.  Tiles are dealt round-robin (`tile_first`, `tile_step`), which is the rank sharding of test-view
// rendering: rank r of n calls it with tile_first = r, tile_step = n and owns exactly its tiles; no collective.
#include "common.cuh"
#include "gemm.h"

namespace rn {

static size_t rup(size_t x) { return (x + 255) / 256 * 256; }

struct RenderWs {
  float *ro, *rd, *vd, *z_c, *pts_c, *raw_c, *w_c, *rgb_c, *depth_c, *acc_c, *z_f, *pts_f, *raw_f, *w_f;
  void* mlp;
  size_t total;
};

static RenderWs carve_render(void* base, int64_t T, int Nc, int Nf) {
  RenderWs w{};
  uint8_t* p = reinterpret_cast<uint8_t*>(rup(reinterpret_cast<uintptr_t>(base)));
  uint8_t* p0 = p;
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(p); p += rup(floats * sizeof(float)); return r; };
  const int Nt = Nc + Nf;
  w.ro = take((size_t)T * 3); w.rd = take((size_t)T * 3); w.vd = take((size_t)T * 3);
  w.z_c = take((size_t)T * Nc); w.pts_c = take((size_t)T * Nc * 3); w.raw_c = take((size_t)T * Nc * 4); w.w_c = take((size_t)T * Nc);
  w.rgb_c = take((size_t)T * 3); w.depth_c = take(T); w.acc_c = take(T);
  if (Nf > 0) {
    w.z_f = take((size_t)T * Nt); w.pts_f = take((size_t)T * Nt * 3); w.raw_f = take((size_t)T * Nt * 4); w.w_f = take((size_t)T * Nt);
  }
  w.mlp = p;
  const size_t m1 = mlp_infer_workspace_bytes(T * Nc, Nc), m2 = Nf > 0 ? mlp_infer_workspace_bytes(T * Nt, Nt) : 0;
  p += rup(m1 > m2 ? m1 : m2);
  w.total = (size_t)(p - p0) + 256;
  return w;
}

}  // namespace rn

using namespace rn;

#define RN_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != RN_OK) return _rc; \
  } while (0)

extern "C" {

int rn_view_dirs(const float* rays_d, int64_t B, float* viewdirs, rn_stream_t stream) {
  if (B == 0) return RN_OK;
  RN_REQUIRE(rays_d && viewdirs && B > 0);
  return launch_view_rays(nullptr, 1, 1.f, 0.f, 0.f, 0, B, nullptr, const_cast<float*>(rays_d), viewdirs, (cudaStream_t)stream);
}

size_t rn_render_workspace_bytes(int64_t tile_rays, int Nc, int Nf) {
  if (tile_rays <= 0 || Nc < 3 || Nf < 0) return 0;
  return carve_render(nullptr, tile_rays, Nc, Nf).total;
}

int rn_render_view(const void* packed_coarse, const void* packed_fine, const float* pose, const float* rays_o_in,
                   const float* rays_d_in, int H, int W, float focal, float cx, float cy, int64_t ray_begin, int64_t ray_end,
                   int64_t tile_rays, int tile_first, int tile_step, const float* z_base, int Nc, const float* u_det, int Nf,
                   int white_background, void* workspace, float* rgb_out, float* depth_out, float* acc_out,
                   int64_t* rays_rendered_host, rn_stream_t stream) {
  RN_REQUIRE(packed_coarse && workspace && rgb_out && z_base && Nc >= 3 && tile_rays > 0 && tile_step >= 1 && tile_first >= 0);
  RN_REQUIRE((pose != nullptr) != (rays_o_in != nullptr && rays_d_in != nullptr));
  RN_REQUIRE(ray_begin >= 0 && ray_end >= ray_begin && (!pose || ray_end <= (int64_t)H * W));
  const bool fine = packed_fine != nullptr && Nf > 0;
  RN_REQUIRE(!fine || u_det);
  cudaStream_t st = (cudaStream_t)stream;
  const RenderWs w = carve_render(workspace, tile_rays, Nc, fine ? Nf : 0);
  const int Nt = Nc + Nf;
  int64_t done = 0, tile = 0;
  for (int64_t a = ray_begin; a < ray_end; a += tile_rays, ++tile) {
    if (tile < tile_first || (tile - tile_first) % tile_step != 0) continue;
    const int64_t B = (ray_end - a < tile_rays) ? ray_end - a : tile_rays;
    const int64_t o = a - ray_begin;                       // offset of this tile in the output / given-ray arrays
    const float* ro = w.ro; const float* rd = w.rd;
    if (pose) {
      RN_TRY(launch_view_rays(pose, W, focal, cx, cy, a, B, w.ro, w.rd, w.vd, st));
    } else {
      ro = rays_o_in + o * 3; rd = rays_d_in + o * 3;
      // view directions of the given rays (with no pose the kernel only READS rd)
      RN_TRY(launch_view_rays(nullptr, W, focal, cx, cy, 0, B, nullptr, const_cast<float*>(rd), w.vd, st));
    }
    // coarse pass: stratified depths without perturbation (rays.py:185-208), coarse network, compositing
    RN_TRY(rn_stratified_fwd(ro, rd, B, z_base, Nc, nullptr, w.z_c, w.pts_c, stream));
    RN_TRY(mlp_infer(packed_coarse, w.pts_c, w.vd, B * Nc, Nc, w.mlp, w.raw_c, st));
    float* rgb_c = fine ? w.rgb_c : rgb_out + o * 3;
    float* dep_c = fine ? w.depth_c : (depth_out ? depth_out + o : w.depth_c);
    float* acc_c = fine ? w.acc_c : (acc_out ? acc_out + o : w.acc_c);
    RN_TRY(rn_composite_fwd(nullptr, nullptr, w.raw_c, w.z_c, rd, nullptr, B, Nc, white_background, 0.0f, rgb_c, dep_c, acc_c,
                            w.w_c, stream));
    if (fine) {
      // inverse-CDF resampling with deterministic draws (rays.py:252: the shared linspace row), fine network on all Nc + Nf
      RN_TRY(rn_sample_hierarchical_fwd(ro, rd, w.z_c, w.w_c, B, Nc, u_det, 0, Nf, w.z_f, w.pts_f, nullptr, stream));
      RN_TRY(mlp_infer(packed_fine, w.pts_f, w.vd, B * Nt, Nt, w.mlp, w.raw_f, st));
      RN_TRY(rn_composite_fwd(nullptr, nullptr, w.raw_f, w.z_f, rd, nullptr, B, Nt, white_background, 0.0f, rgb_out + o * 3,
                              depth_out ? depth_out + o : w.depth_c, acc_out ? acc_out + o : w.acc_c, w.w_f, stream));
    }
    done += B;
  }
  if (rays_rendered_host) *rays_rendered_host = done;
  return RN_OK;
}

}  // extern "C"
