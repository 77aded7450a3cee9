// api.cu — the C ABI of libgftorf_b200.so (include/gftorf.h): argument checks, private workspace
// layout, and the launch sequence of forward / backward / markVisible / dist2.
//
// Stands where rasterize_points.cu:35-304 + CudaRasterizer::Rasterizer::{forward,backward,
// markVisible} (cuda_rasterizer/rasterizer_impl.cu:143-159,215-378,382-499) and spatial.cu:15-26
// stand in the reference.  No torch, no globals (a thread-local error string only), every launch
// on the caller's stream.
//
// Forward launch sequence (reference: 1 + CUB scan + D2H + 1 + CUB sort (8 launches) + memset + 1
// + 1, preceded by 11 torch fills):
//   memset(scan header) -> preprocess_fwd (cull, project, SH, phasor, tile rect, chained scan,
//   zero pixels/ranges) -> D2H R -> duplicate_keys -> histogram + N onesweep passes ->
//   identify_ranges -> blend_fwd (writes every output plane, including the always-zero ones)
// Backward: memset(grad records) -> blend_bwd -> preprocess_bwd (cov2D + SH + phasor + cov3D
// backward, writes every gradient row once; reference: 15 torch memsets + 3 kernels).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/gftorf.h"
#include "common.cuh"
#include "kernels.h"
#include "radix_sort.cuh"

#include <atomic>
#include <vector>

namespace gft {
static std::atomic<unsigned long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace gft

namespace {

thread_local std::string g_err;

// Optional per-stage timing with CUDA events on the caller's stream (gft_profile_* in the ABI).
struct StageRec { const char* name; cudaEvent_t e0, e1; };
struct Profile {
  bool on = false;
  std::vector<StageRec> pool;   // events are created once and reused
  int used = 0;
  void reset() { used = 0; }
  StageRec* next(const char* name) {
    if (used == (int)pool.size()) {
      StageRec r; r.name = name;
      cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
      pool.push_back(r);
    }
    pool[used].name = name;
    return &pool[used++];
  }
};
thread_local Profile g_prof;

struct Stage {
  StageRec* r = nullptr;
  cudaStream_t s;
  Stage(const char* name, cudaStream_t stream) : s(stream) {
    if (g_prof.on) { r = g_prof.next(name); cudaEventRecord(r->e0, s); }
  }
  ~Stage() { if (r) cudaEventRecord(r->e1, s); }
};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
}  // namespace
namespace gft {
int set_error(int code, const char* msg) { return fail(code, msg); }   // for the other .cu files
}  // namespace gft
namespace {

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

// Number of key bits that hold the tile id: restates getHigherMsb (rasterizer_impl.cu:35-50),
// which returns the bit length of n (1 for n == 0).
int tile_bits(uint32_t n) {
  int b = 0;
  while (b < 32 && (n >> b) != 0u) ++b;
  return b == 0 ? 1 : b;
}

struct GeomWs {
  size_t scan_hdr;  // [ticket u32][num_rendered u32][pad to 16][state u64[blocks]]
  size_t scan_hdr_bytes;
  size_t rec, depths, tiles_touched, point_offsets, rect, cov3D, clamped, pa, total;
};

GeomWs geom_layout(int P) {
  GeomWs g;
  const size_t n = (size_t)(P > 0 ? P : 1);
  const size_t blocks = (n + GFT_BLOCK - 1) / GFT_BLOCK;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
  g.scan_hdr_bytes = 16 + blocks * 8;
  g.scan_hdr = take(g.scan_hdr_bytes);
  g.rec = take(n * GFT_REC_FLOATS * 4);
  g.depths = take(n * 4);
  g.tiles_touched = take(n * 4);
  g.point_offsets = take(n * 4);
  g.rect = take(n * 8);
  g.cov3D = take(n * 24);
  g.clamped = take(n * 4);
  g.pa = take(n * 8);
  g.total = off;
  return g;
}

struct ImgWs {
  size_t state, ranges, total;
};

ImgWs img_layout(int W, int H) {
  ImgWs m;
  const size_t N = (size_t)W * H;
  const size_t T = (size_t)((W + GFT_TILE_X - 1) / GFT_TILE_X) * ((H + GFT_TILE_Y - 1) / GFT_TILE_Y);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
  m.state = take((N > 0 ? N : 1) * 16);
  m.ranges = take((T > 0 ? T : 1) * 8);
  m.total = off;
  return m;
}

struct BinWs {
  size_t keys_a, keys_b, vals_a, vals_b, temp, temp_bytes, total;
};

BinWs bin_layout(int R) {
  BinWs b;
  const size_t n = (size_t)(R > 0 ? R : 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
  b.vals_a = take(n * 4);   // offset 0: where the sorted point list always ends up
  b.vals_b = take(n * 4);
  b.keys_a = take(n * 8);
  b.keys_b = take(n * 8);
  b.temp_bytes = gft::radix_sort_temp_bytes((int)n);
  b.temp = take(b.temp_bytes);
  b.total = off;
  return b;
}

// Pinned word + event per device for reading num_rendered back without a full stream sync.
struct HostSync { cudaEvent_t ev; uint32_t* pinned; };
thread_local HostSync g_hs[64] = {};
HostSync* host_sync() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HostSync& h = g_hs[dev];
  if (!h.pinned) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&h.pinned), 64, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&h.ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &h;
}

#define GFT_CUDA_OK(stage)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ == cudaSuccess && debug) e__ = cudaStreamSynchronize(stream);                   \
    if (e__ != cudaSuccess)                                                                 \
      return fail(-2, std::string("CUDA error in ") + stage + ": " + cudaGetErrorString(e__)); \
  } while (0)

}  // namespace

extern "C" {

const char* gft_last_error(void) { return g_err.c_str(); }

unsigned long long gft_launch_count(void) { return gft::g_launches.load(); }
void gft_profile_enable(int on) { g_prof.on = on != 0; g_prof.reset(); }
int gft_profile_read(float* ms, const char** names, int cap) {
  int n = 0;
  for (int i = 0; i < g_prof.used && n < cap; ++i) {
    StageRec& r = g_prof.pool[i];
    if (cudaEventSynchronize(r.e1) != cudaSuccess) break;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) break;
    ms[n] = t;
    if (names) names[n] = r.name;
    ++n;
  }
  g_prof.reset();
  return n;
}
int gft_abi_version(void) { return GFT_ABI_VERSION; }

size_t gft_geom_bytes(int P) { return geom_layout(P).total; }
size_t gft_img_bytes(int width, int height) { return img_layout(width, height).total; }
size_t gft_binning_bytes(int R) { return bin_layout(R).total; }
size_t gft_backward_scratch_bytes(int P) {
  return up256((size_t)(P > 0 ? P : 1) * GFT_GRAD_FLOATS * 4);
}

void gft_workspace_layout(int P, int R, int width, int height, GftWorkspaceLayout* o) {
  std::memset(o, 0, sizeof(*o));
  const GeomWs g = geom_layout(P);
  o->geom_rec = g.rec;
  o->geom_depths = g.depths;
  o->geom_tiles_touched = g.tiles_touched;
  o->geom_point_offsets = g.point_offsets;
  o->geom_rect = g.rect;
  o->geom_cov3D = g.cov3D;
  o->geom_clamped = g.clamped;
  o->geom_pa = g.pa;
  o->geom_total = g.total;
  const BinWs b = bin_layout(R);
  const int gx = (width + GFT_TILE_X - 1) / GFT_TILE_X, gy = (height + GFT_TILE_Y - 1) / GFT_TILE_Y;
  (void)gx; (void)gy;
  o->bin_keys = b.keys_a;             // the ping-pong is started so that it always ends in "a"
  o->bin_keys_unsorted = b.keys_b;    // scratch: overwritten by the passes
  o->bin_point_list = b.vals_a;
  o->bin_point_list_unsorted = b.vals_b;
  o->bin_total = b.total;
  const ImgWs m = img_layout(width, height);
  o->img_state = m.state;
  o->img_ranges = m.ranges;
  o->img_total = m.total;
}

int gft_forward(const GftForwardArgs* a, gft_alloc_fn geom_alloc, gft_alloc_fn binning_alloc,
                gft_alloc_fn img_alloc, void* ctx, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(-1, "gft_forward: null args");
  const bool debug = a->debug != 0;
  const int P = a->P, W = a->width, H = a->height;
  if (P < 0 || W <= 0 || H <= 0) return fail(-1, "gft_forward: bad P / width / height");
  if (a->sh_degree < 0 || a->sh_degree > 3) return fail(-1, "gft_forward: sh_degree must be 0..3");
  if (!a->out_color || !a->out_phasor || !a->out_depth || !a->out_acc || !a->out_depth_distortion ||
      !a->out_distribution || !a->radii || !a->pixels || !a->background)
    return fail(-1, "gft_forward: required output/background pointer is null");
  const size_t N = (size_t)W * H;

  if (P == 0) {
    // rasterize_points.cu:104 — nothing is rendered, outputs stay at their zero fill, R = 0
    cudaMemsetAsync(a->out_color, 0, 3 * N * 4, stream);
    cudaMemsetAsync(a->out_phasor, 0, 7 * N * 4, stream);
    cudaMemsetAsync(a->out_depth, 0, N * 4, stream);
    cudaMemsetAsync(a->out_acc, 0, N * 4, stream);
    cudaMemsetAsync(a->out_depth_distortion, 0, N * 4, stream);
    cudaMemsetAsync(a->out_distribution, 0, 3 * N * 4, stream);
    if (a->out_normal) cudaMemsetAsync(a->out_normal, 0, 3 * N * 4, stream);
    if (a->out_entropy) cudaMemsetAsync(a->out_entropy, 0, N * 4, stream);
    if (a->out_amp_distortion) cudaMemsetAsync(a->out_amp_distortion, 0, N * 4, stream);
    GFT_CUDA_OK("forward(P=0)");
    return 0;
  }
  if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->campos)
    return fail(-1, "gft_forward: required input pointer is null");
  if (!a->cov3D_precomp && (!a->scales || !a->rotations))
    return fail(-1, "gft_forward: need scales+rotations or cov3D_precomp");
  // rasterizer_impl.cu:269-272 (NUM_CHANNELS is 3, so this cannot fire; colours may be absent
  // only on paths that do not read them)
  if (a->shs && a->M < (a->sh_degree + 1) * (a->sh_degree + 1))
    return fail(-1, "gft_forward: shs has fewer coefficients than the active degree needs");
  if (a->shs_p && a->M_p < (a->sh_degree + 1) * (a->sh_degree + 1))
    return fail(-1, "gft_forward: shs_p has fewer coefficients than the active degree needs");

  const int gx = (W + GFT_TILE_X - 1) / GFT_TILE_X, gy = (H + GFT_TILE_Y - 1) / GFT_TILE_Y;
  const GeomWs gl = geom_layout(P);
  const ImgWs il = img_layout(W, H);
  char* geom = geom_alloc(ctx, gl.total);
  char* img = img_alloc(ctx, il.total);
  if (!geom || !img) return fail(-3, "gft_forward: workspace callback returned null");

  uint32_t* hdr = reinterpret_cast<uint32_t*>(geom + gl.scan_hdr);
  cudaMemsetAsync(hdr, 0, gl.scan_hdr_bytes, stream);

  gft::PreprocessParams pp;
  std::memset(&pp, 0, sizeof(pp));
  pp.P = P; pp.D = a->sh_degree; pp.M = a->M; pp.M_p = a->M_p;
  pp.W = W; pp.H = H; pp.grid_x = gx; pp.grid_y = gy; pp.num_tiles = gx * gy;
  pp.means3D = a->means3D; pp.scales = a->scales; pp.scale_modifier = a->scale_modifier;
  pp.rotations = a->rotations; pp.opacities = a->opacities; pp.shs = a->shs; pp.shs_p = a->shs_p;
  pp.cov3D_precomp = a->cov3D_precomp; pp.colors_precomp = a->colors_precomp;
  pp.phasors_precomp = a->phasors_precomp; pp.viewmatrix = a->viewmatrix;
  pp.projmatrix = a->projmatrix; pp.campos = a->campos;
  pp.tan_fovx = a->tan_fovx; pp.tan_fovy = a->tan_fovy;
  pp.focal_y = H / (2.0f * a->tan_fovy);  // rasterizer_impl.cu:249-250
  pp.focal_x = W / (2.0f * a->tan_fovx);
  pp.prefiltered = a->prefiltered; pp.near_n = a->near_n; pp.far_n = a->far_n;
  pp.dist2phase = 4.0f * GFT_PI_F / a->depth_range;  // forward.cu:752
  pp.use_view_dependent_phase = a->use_view_dependent_phase;
  pp.phase_offset = a->phase_offset; pp.dc_offset = a->dc_offset;
  pp.radii = a->radii; pp.pixels = a->pixels;
  pp.rec = reinterpret_cast<float*>(geom + gl.rec);
  pp.depths = reinterpret_cast<float*>(geom + gl.depths);
  pp.tiles_touched = reinterpret_cast<uint32_t*>(geom + gl.tiles_touched);
  pp.point_offsets = reinterpret_cast<uint32_t*>(geom + gl.point_offsets);
  pp.rect = reinterpret_cast<uint16_t*>(geom + gl.rect);
  pp.cov3D = reinterpret_cast<float*>(geom + gl.cov3D);
  pp.clamped = reinterpret_cast<uint32_t*>(geom + gl.clamped);
  pp.pa = reinterpret_cast<float*>(geom + gl.pa);
  pp.ranges = reinterpret_cast<uint2*>(img + il.ranges);
  pp.scan_ticket = hdr;
  const gft::KeyFormat kf = gft::key_format(a->near_n, a->far_n);
  pp.key_format_out = hdr + 2;
  pp.key_depth_bits = kf.depth_bits; pp.key_depth_base = kf.depth_base;
  pp.num_rendered = hdr + 1;
  pp.scan_state = reinterpret_cast<unsigned long long*>(hdr + 4);
  const char* nocull = std::getenv("GFT_NO_CULL");
  pp.subtile_cull = !(nocull && nocull[0] == '1');
  { Stage st("preprocess_fwd", stream); gft::launch_preprocess_fwd(pp, stream); }
  GFT_CUDA_OK("preprocess");

  // R sizes the binning workspace, so it has to reach the host (rasterizer_impl.cu:310-315).  It
  // travels through a pinned word and an event, so the host waits for the preprocess kernel only.
  HostSync* hs = host_sync();
  if (!hs) return fail(-2, "gft_forward: cannot create the pinned word / event for num_rendered");
  uint32_t* d_R = hdr + 1;
  cudaError_t e = cudaMemcpyAsync(hs->pinned, d_R, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaEventRecord(hs->ev, stream);
  if (e != cudaSuccess)
    return fail(-2, std::string("CUDA error reading num_rendered: ") + cudaGetErrorString(e));

  const int end_bit = kf.depth_bits + tile_bits((uint32_t)(gx * gy));
  uint2* ranges = pp.ranges;

  // Everything after the preprocess kernel, for a binning workspace of `cap` pairs.  With
  // dev_count != nullptr the kernels read the actual pair count from device memory, so they can be
  // enqueued before the host knows it.
  auto bin_and_blend = [&](int cap, const uint32_t* dev_count) -> int {
    const BinWs bl = bin_layout(cap);
    char* bin = binning_alloc(ctx, bl.total);
    if (!bin) return fail(-3, "gft_forward: binning workspace callback returned null");
    uint64_t* keys_a = reinterpret_cast<uint64_t*>(bin + bl.keys_a);
    uint64_t* keys_b = reinterpret_cast<uint64_t*>(bin + bl.keys_b);
    uint32_t* vals_a = reinterpret_cast<uint32_t*>(bin + bl.vals_a);
    uint32_t* vals_b = reinterpret_cast<uint32_t*>(bin + bl.vals_b);
    if (cap > 0) {
      // start the ping-pong in the buffer that makes the last pass land in (keys_a, vals_a)
      const bool flip = gft::sort_result_in_out(end_bit);
      uint64_t* kin = flip ? keys_b : keys_a; uint64_t* kout = flip ? keys_a : keys_b;
      uint32_t* vin = flip ? vals_b : vals_a; uint32_t* vout = flip ? vals_a : vals_b;
      { Stage st("duplicate_keys", stream);
        gft::launch_duplicate_keys(P, a->radii, pp.rect, pp.depths, pp.point_offsets, kin, vin, gx,
                                   (uint32_t)cap, kf, stream); }
      GFT_CUDA_OK("duplicate_keys");
      int rc;
      { Stage st("radix_sort", stream);
        rc = gft::sort_pairs(bin + bl.temp, bl.temp_bytes, kin, kout, vin, vout, cap, end_bit, stream,
                             dev_count); }
      if (rc < 0) return fail(-2, "gft_forward: radix sort failed");
      GFT_CUDA_OK("sort");
      { Stage st("identify_ranges", stream);
        gft::launch_identify_ranges(cap, dev_count, keys_a, ranges, kf, stream); }
      GFT_CUDA_OK("identify_ranges");
    }
    gft::BlendFwdParams bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.W = W; bp.H = H; bp.grid_x = gx; bp.grid_y = gy;
    bp.ranges = ranges; bp.point_list = vals_a;
    bp.rec = reinterpret_cast<const float4*>(pp.rec);
    bp.bg = a->background; bp.bg_mode = a->bg_mode;
    bp.img_state = reinterpret_cast<float4*>(img + il.state);
    bp.out_color = a->out_color; bp.out_phasor = a->out_phasor; bp.out_depth = a->out_depth;
    bp.out_normal = a->out_normal; bp.out_acc = a->out_acc; bp.out_entropy = a->out_entropy;
    bp.out_depth_distortion = a->out_depth_distortion;
    bp.out_amp_distortion = a->out_amp_distortion; bp.out_distribution = a->out_distribution;
    bp.pixels = a->pixels;
    { Stage st("blend_fwd", stream); gft::launch_blend_fwd(bp, stream); }
    GFT_CUDA_OK("blend_fwd");
    return 0;
  };
  auto wait_R = [&](int& R_out) -> int {
    const cudaError_t ee = cudaEventSynchronize(hs->ev);
    if (ee != cudaSuccess)
      return fail(-2, std::string("CUDA error reading num_rendered: ") + cudaGetErrorString(ee));
    const uint32_t R_u = *hs->pinned;
    if (R_u > 0x7fffffffu) return fail(-4, "gft_forward: num_rendered exceeds 2^31-1");
    R_out = (int)R_u;
    return 0;
  };

  int R = 0;
  const bool hinted = a->R_hint > 0 && gft::sort_backend() == 0;
  if (!hinted) {
    // exact mode (reference behaviour): wait for the count, then size the workspace from it
    int rc = wait_R(R);
    if (rc < 0) return rc;
    rc = bin_and_blend(R, nullptr);
    if (rc < 0) return rc;
  } else {
    // hinted mode: the caller's estimate sizes the workspace, all kernels are enqueued at once and
    // read the count on the device; the host only learns R afterwards (no bubble in the stream).
    int rc = bin_and_blend(a->R_hint, d_R);
    if (rc < 0) return rc;
    rc = wait_R(R);
    if (rc < 0) return rc;
    if (R > a->R_hint) {   // estimate too small: redo binning + blend with the exact size
      cudaMemsetAsync(a->pixels, 0, (size_t)P * sizeof(float), stream);
      cudaMemsetAsync(ranges, 0, (size_t)gx * gy * sizeof(uint2), stream);
      rc = bin_and_blend(R, nullptr);
      if (rc < 0) return rc;
    }
  }
  return R;
}

int gft_backward(const GftBackwardArgs* a, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(-1, "gft_backward: null args");
  const bool debug = a->debug != 0;
  const int P = a->P, W = a->width, H = a->height;
  if (P < 0 || W <= 0 || H <= 0) return fail(-1, "gft_backward: bad P / width / height");
  if (!a->dL_dphase_offset || !a->dL_ddc_offset)
    return fail(-1, "gft_backward: dL_dphase_offset / dL_ddc_offset are required");
  if (P == 0) {  // rasterize_points.cu:238
    if (!a->accumulate) {
      cudaMemsetAsync(a->dL_dphase_offset, 0, 4, stream);
      cudaMemsetAsync(a->dL_ddc_offset, 0, 4, stream);
    }
    GFT_CUDA_OK("backward(P=0)");
    return 0;
  }
  if (!a->geom_buffer || !a->img_buffer || !a->radii || !a->means3D || !a->viewmatrix ||
      !a->projmatrix || !a->campos || !a->background)
    return fail(-1, "gft_backward: required input pointer is null");
  if (a->R > 0 && !a->binning_buffer) return fail(-1, "gft_backward: binning buffer is null");
  if (!a->dL_dout_color || !a->dL_dout_phasor || !a->dL_dout_depth || !a->dL_dout_acc ||
      !a->dL_dout_depth_distortion)
    return fail(-1, "gft_backward: incoming gradient pointer is null");
  if (!a->dL_dmeans2D || !a->dL_dopacity || !a->dL_dmeans3D || !a->scratch)
    return fail(-1, "gft_backward: required output pointer is null");
  if (a->dL_dphasors)
    return fail(-1, "gft_backward: dL_dphasors is a reference-only field (the 7 phasor gradients are "
                    "reduced to the 4 combinations the phasor backward consumes); pass NULL");
  if ((a->shs && !a->dL_dsh) || (a->shs_p && !a->dL_dsh_p) ||
      (a->scales && (!a->dL_dscales || !a->dL_drotations)))
    return fail(-1, "gft_backward: missing gradient buffer for a provided input");

  const int gx = (W + GFT_TILE_X - 1) / GFT_TILE_X, gy = (H + GFT_TILE_Y - 1) / GFT_TILE_Y;
  const GeomWs gl = geom_layout(P);
  const ImgWs il = img_layout(W, H);
  const char* geom = a->geom_buffer;
  const char* img = a->img_buffer;
  const char* bin = a->binning_buffer;

  { Stage st("zero_grad_records", stream);
    cudaMemsetAsync(a->scratch, 0, (size_t)P * GFT_GRAD_FLOATS * 4, stream); }

  if (a->R > 0) {
    gft::BlendBwdParams bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.W = W; bp.H = H; bp.grid_x = gx; bp.grid_y = gy;
    bp.ranges = reinterpret_cast<const uint2*>(img + il.ranges);
    bp.point_list = reinterpret_cast<const uint32_t*>(bin);  // sorted list: offset 0 of the workspace
    bp.rec = reinterpret_cast<const float4*>(geom + gl.rec);
    bp.bg = a->background; bp.bg_mode = a->bg_mode;
    bp.img_state = reinterpret_cast<const float4*>(img + il.state);
    bp.dL_dcolor = a->dL_dout_color; bp.dL_dphasor = a->dL_dout_phasor;
    bp.dL_ddepth = a->dL_dout_depth; bp.dL_dacc = a->dL_dout_acc;
    bp.dL_ddd = a->dL_dout_depth_distortion;
    bp.grad_rec = a->scratch;
    { Stage st("blend_bwd", stream); gft::launch_blend_bwd(bp, stream); }
    GFT_CUDA_OK("blend_bwd");
  }

  gft::PreprocessBwdParams pb;
  std::memset(&pb, 0, sizeof(pb));
  pb.P = P; pb.D = a->sh_degree; pb.M = a->M; pb.M_p = a->M_p; pb.W = W; pb.H = H;
  pb.means3D = a->means3D; pb.radii = a->radii; pb.shs = a->shs; pb.shs_p = a->shs_p;
  pb.clamped = reinterpret_cast<const uint32_t*>(geom + gl.clamped);
  pb.scales = a->scales; pb.rotations = a->rotations; pb.scale_modifier = a->scale_modifier;
  // rasterizer_impl.cu:471
  pb.cov3D = a->cov3D_precomp ? a->cov3D_precomp : reinterpret_cast<const float*>(geom + gl.cov3D);
  pb.viewmatrix = a->viewmatrix; pb.projmatrix = a->projmatrix; pb.campos = a->campos;
  pb.focal_y = H / (2.0f * a->tan_fovy);
  pb.focal_x = W / (2.0f * a->tan_fovx);
  pb.tan_fovx = a->tan_fovx; pb.tan_fovy = a->tan_fovy;
  pb.rec = reinterpret_cast<const float*>(geom + gl.rec);
  pb.pa = reinterpret_cast<const float*>(geom + gl.pa);
  pb.grad_rec = a->scratch;
  pb.near_n = a->near_n; pb.far_n = a->far_n;
  pb.dist2phase = 4.0f * GFT_PI_F / a->depth_range;  // backward.cu:936
  pb.use_view_dependent_phase = a->use_view_dependent_phase;
  pb.phase_offset = a->phase_offset; pb.dc_offset = a->dc_offset;
  pb.accumulate = a->accumulate;
  pb.dL_dmeans2D = a->dL_dmeans2D; pb.dL_dopacity = a->dL_dopacity;
  pb.dL_dmeans3D = a->dL_dmeans3D; pb.dL_dsh = a->dL_dsh; pb.dL_dsh_p = a->dL_dsh_p;
  pb.dL_dscales = a->dL_dscales; pb.dL_drotations = a->dL_drotations;
  pb.dL_dphase_offset = a->dL_dphase_offset; pb.dL_ddc_offset = a->dL_ddc_offset;
  pb.dL_dcolors = a->dL_dcolors; pb.dL_dcov3D = a->dL_dcov3D;
  pb.dL_dconic = a->dL_dconic; pb.dL_ddist = a->dL_ddist; pb.dL_dndc = a->dL_dndc;
  { Stage st("preprocess_bwd", stream); gft::launch_preprocess_bwd(pb, stream); }
  GFT_CUDA_OK("preprocess_bwd");
  return 0;
}

int gft_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* /*projmatrix*/, uint8_t* present, float near_n, float far_n,
                     gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P < 0) return fail(-1, "gft_mark_visible: bad P");
  if (P == 0) return 0;
  if (!means3D || !viewmatrix || !present) return fail(-1, "gft_mark_visible: null pointer");
  gft::launch_mark_visible(P, means3D, viewmatrix, present, near_n, far_n, stream);
  GFT_CUDA_OK("mark_visible");
  return 0;
}

size_t gft_dist2_workspace_bytes(int P) { return gft::knn_workspace_bytes(P); }

int gft_dist2(const float* points, int P, float* out, char* workspace, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P < 0) return fail(-1, "gft_dist2: bad P");
  if (P == 0) return 0;
  if (!points || !out || !workspace) return fail(-1, "gft_dist2: null pointer");
  const int rc = gft::knn_dist2(points, P, out, workspace, stream);
  if (rc < 0) return fail(-2, "gft_dist2: radix sort failed");
  GFT_CUDA_OK("dist2");
  return 0;
}

}  // extern "C"
