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

// ---- private workspace layouts ------------------------------------------------------------------
// geometry: cov3D [P][6] (view independent) then view-major arrays [V][P][...]
struct GeomWs {
  size_t cov3D, rec, depths, tiles_touched, rect, clamped, pa, total;
};

GeomWs geom_layout(int P, int V) {
  GeomWs g;
  const size_t n = (size_t)(P > 0 ? P : 1), nv = n * (size_t)(V > 0 ? V : 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
  g.cov3D = take(n * 24);
  g.rec = take(nv * GFT_REC_FLOATS * 4);
  g.depths = take(nv * 4);
  g.tiles_touched = take(nv * 4);
  g.rect = take(nv * 8);
  g.clamped = take(nv * 4);
  g.pa = take(nv * 8);
  g.total = off;
  return g;
}

// image: header (word 0 = scan ticket, word 1 = num_rendered, bytes 256.. = scan chain state) |
// sub-bin counters | cursors | starts | tile ranges | launch order | pixel state
struct ImgWs {
  size_t hdr, counts, cursors, starts, ranges, order, state, total;
  size_t T_total, N_total;
  int S;     // sub-counters per tile
};

// Sub-counters per tile: as many as the option allows while one block can still scan them all.
int sub_bins_for(size_t T_total) {
  int S = gft::option(gft::OPT_SUB_BINS);
  if (S < 1) S = 1;
  if (S > 16) S = 16;
  while (S & (S - 1)) S &= S - 1;            // power of two
  while (S > 1 && T_total * (size_t)S > 131072) S >>= 1;
  return S;
}

ImgWs img_layout(size_t T_total, size_t N_total) {
  ImgWs m;
  m.T_total = T_total; m.N_total = N_total;
  m.S = sub_bins_for(T_total);
  const size_t nb = (T_total > 0 ? T_total : 1) * (size_t)m.S;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
  m.hdr = take(512);       // 16 header bytes + 32 x 8 B of scan state (n <= 131072 -> <= 32 scan blocks)
  m.counts = take(nb * 4);
  m.cursors = take(nb * 4);
  m.starts = take((nb + 1) * 4);
  m.ranges = take((T_total > 0 ? T_total : 1) * 8);
  m.order = take((T_total > 0 ? T_total : 1) * 4);
  m.state = take((N_total > 0 ? N_total : 1) * 16);
  m.total = off;
  return m;
}

// binning: sorted Gaussian ids (offset 0) | 64-bit entries (depth bits << 32 | id), sorted in place
struct BinWs {
  size_t point_list, entries, total;
};

BinWs bin_layout(int R) {
  BinWs b;
  const size_t n = (size_t)(R > 0 ? R : 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
  b.point_list = take(n * 4);
  b.entries = take(n * 8);
  b.total = off;
  return b;
}

struct ViewDims { int gx, gy, tile_base; size_t pix_base; };

// Mapped pinned word + event per device: the scan kernel stores num_rendered straight into host
// memory, the host waits for the event recorded behind it.  No stream sync and no copy-engine
// transfer (a 4-byte cudaMemcpyAsync queues behind any bulk D2H copy the application has in
// flight on another stream — measured: 3 ms per forward in a pipelined training loop).
struct HostSync { cudaEvent_t ev; uint32_t* pinned; uint32_t* pinned_dev; };
thread_local HostSync g_hs[64] = {};
HostSync* host_sync() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HostSync& h = g_hs[dev];
  if (!h.pinned) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&h.pinned), 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&h.pinned_dev), h.pinned, 0) != cudaSuccess) return nullptr;
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

bool subtile_cull_enabled() { return gft::option(gft::OPT_NO_CULL) == 0; }

void fill_cam(gft::ViewCam& c, const GftViewArgs& v, const ViewDims& d) {
  c.W = v.width; c.H = v.height; c.grid_x = d.gx; c.grid_y = d.gy; c.tile_base = d.tile_base;
  c.use_view_dependent_phase = v.use_view_dependent_phase;
  c.tan_fovx = v.tan_fovx; c.tan_fovy = v.tan_fovy;
  c.focal_y = v.height / (2.0f * v.tan_fovy);  // rasterizer_impl.cu:249-250
  c.focal_x = v.width / (2.0f * v.tan_fovx);
  c.near_n = v.near_n; c.far_n = v.far_n;
  c.dist2phase = 4.0f * GFT_PI_F / v.depth_range;  // forward.cu:752, backward.cu:936
  c.phase_offset = v.phase_offset; c.dc_offset = v.dc_offset;
  c.pad_ = 0;
  c.viewmatrix = v.viewmatrix; c.projmatrix = v.projmatrix; c.campos = v.campos;
  c.radii = v.radii; c.pixels = v.pixels; c.dL_dmeans2D = v.dL_dmeans2D;
}

// tile / pixel bases of the views of a batch; returns false on bad sizes
bool view_dims(const GftViewArgs* views, int n, ViewDims* d, size_t& T_total, size_t& N_total) {
  T_total = 0; N_total = 0;
  for (int i = 0; i < n; ++i) {
    if (views[i].width <= 0 || views[i].height <= 0) return false;
    d[i].gx = (views[i].width + GFT_TILE_X - 1) / GFT_TILE_X;
    d[i].gy = (views[i].height + GFT_TILE_Y - 1) / GFT_TILE_Y;
    d[i].tile_base = (int)T_total;
    d[i].pix_base = N_total;
    T_total += (size_t)d[i].gx * d[i].gy;
    N_total += (size_t)views[i].width * views[i].height;
  }
  return T_total < 0x7fffffffull;
}

}  // namespace

namespace gft {
// Tunables / A-B switches: read from the environment once, changeable through gft_set_option
// (tests and measurement scripts flip them inside one process; no getenv on the call path).
namespace {
struct OptDef { const char* name; const char* env; int def; };
const OptDef kOpts[OPT_COUNT] = {
    {"sort_cap", "GFT_SORT_CAP", 0},     // entries per block of the tile sort held in shared memory (0 = automatic)
    {"bwd_pred", "GFT_BWD_PRED", 1},     // branch-free replay in the blend backward
    {"pbwd_minb", "GFT_PBWD_MINB", 3},   // resident blocks per SM the preprocess backward is compiled for (3 or 4)
    {"no_cull", "GFT_NO_CULL", 0},       // 1: no sub-tile culling extents
    {"sort_radix", "GFT_SORT_RADIX", 1}, // 0: bitonic network instead of the shared-memory radix sort per tile
    {"sub_bins", "GFT_SUB_BINS", 16},    // sub-counters per tile of the binning (power of two, <= 16)
    {"sort_match", "GFT_SORT_MATCH", 0}, // 1: MATCH.ANY instead of eight ballots for the radix ranking
    {"tile_order", "GFT_TILE_ORDER", 1}, // 0: blend blocks in tile index order instead of longest list first
    {"bwd_ring", "GFT_BWD_RING", 1},     // 0: block-wide double buffer instead of the mbarrier ring in the blend backward
    {"pfwd_minb", "GFT_PFWD_MINB", 3},   // resident blocks per SM the preprocess forward is compiled for (3 or 4)
    {"blend_half", "GFT_BLEND_HALF", 0}, // 1: blend warps walk their 8x4 patch as two independent 4x4 halves
    {"sort_adapt", "GFT_SORT_ADAPT", 1}, // 0: four radix passes over all 32 depth bits instead of <= 3 over the bits that vary in the tile
};
std::atomic<int> g_opt[OPT_COUNT];
std::atomic<bool> g_opt_init{false};
void init_options() {
  if (g_opt_init.load(std::memory_order_acquire)) return;
  for (int i = 0; i < OPT_COUNT; ++i) {
    const char* e = std::getenv(kOpts[i].env);
    g_opt[i].store(e ? std::atoi(e) : kOpts[i].def, std::memory_order_relaxed);
  }
  g_opt_init.store(true, std::memory_order_release);
}
}  // namespace
int option(int id) {
  init_options();
  return g_opt[id].load(std::memory_order_relaxed);
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    __atomic_store_n(&cached[dev], n, __ATOMIC_RELAXED);
  }
  return n;
}
}  // namespace gft

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

int gft_set_option(const char* name, int value) {
  gft::init_options();
  for (int i = 0; i < gft::OPT_COUNT; ++i)
    if (name && std::strcmp(name, gft::kOpts[i].name) == 0) return gft::g_opt[i].exchange(value);
  return fail(-1, std::string("gft_set_option: unknown option ") + (name ? name : "(null)"));
}

size_t gft_geom_bytes(int P) { return geom_layout(P, 1).total; }
size_t gft_geom_bytes_views(int P, int n_views) { return geom_layout(P, n_views).total; }
size_t gft_img_bytes(int width, int height) {
  const size_t T = (size_t)((width + GFT_TILE_X - 1) / GFT_TILE_X) * ((height + GFT_TILE_Y - 1) / GFT_TILE_Y);
  return img_layout(T, (size_t)width * height).total;
}
size_t gft_img_bytes_views(int n_views, const int* widths, const int* heights) {
  size_t T = 0, N = 0;
  for (int i = 0; i < n_views; ++i) {
    T += (size_t)((widths[i] + GFT_TILE_X - 1) / GFT_TILE_X) * ((heights[i] + GFT_TILE_Y - 1) / GFT_TILE_Y);
    N += (size_t)widths[i] * heights[i];
  }
  return img_layout(T, N).total;
}
size_t gft_binning_bytes(int R) { return bin_layout(R).total; }
size_t gft_backward_scratch_bytes(int P) {
  return up256((size_t)(P > 0 ? P : 1) * GFT_GRAD_FLOATS * 4);
}
size_t gft_backward_scratch_bytes_views(int P, int n_views) {
  return up256((size_t)(P > 0 ? P : 1) * (size_t)(n_views > 0 ? n_views : 1) * GFT_GRAD_FLOATS * 4);
}

void gft_workspace_layout_views(int P, int R, int n_views, const int* widths, const int* heights,
                                GftWorkspaceLayout* o) {
  std::memset(o, 0, sizeof(*o));
  const GeomWs g = geom_layout(P, n_views);
  o->geom_cov3D = g.cov3D;
  o->geom_rec = g.rec;
  o->geom_depths = g.depths;
  o->geom_tiles_touched = g.tiles_touched;
  o->geom_rect = g.rect;
  o->geom_clamped = g.clamped;
  o->geom_pa = g.pa;
  o->geom_total = g.total;
  const BinWs b = bin_layout(R);
  o->bin_point_list = b.point_list;
  o->bin_entries = b.entries;
  o->bin_total = b.total;
  size_t T = 0, N = 0;
  for (int i = 0; i < n_views; ++i) {
    T += (size_t)((widths[i] + GFT_TILE_X - 1) / GFT_TILE_X) * ((heights[i] + GFT_TILE_Y - 1) / GFT_TILE_Y);
    N += (size_t)widths[i] * heights[i];
  }
  const ImgWs m = img_layout(T, N);
  o->img_hdr = m.hdr;
  o->img_sub_bins = (size_t)m.S;
  o->img_tile_counts = m.counts;
  o->img_ranges = m.ranges;
  o->img_state = m.state;
  o->img_total = m.total;
}

void gft_workspace_layout(int P, int R, int width, int height, GftWorkspaceLayout* o) {
  gft_workspace_layout_views(P, R, 1, &width, &height, o);
}

int gft_forward_views(const GftForwardViewsArgs* a, gft_alloc_fn geom_alloc,
                      gft_alloc_fn binning_alloc, gft_alloc_fn img_alloc, void* ctx,
                      gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(-1, "gft_forward: null args");
  const bool debug = a->debug != 0;
  const int P = a->P, NV = a->n_views;
  if (P < 0) return fail(-1, "gft_forward: bad P");
  if (NV < 1 || NV > GFT_MAX_VIEWS || !a->views)
    return fail(-1, "gft_forward: n_views must be 1.." + std::to_string(GFT_MAX_VIEWS));
  if (a->sh_degree < 0 || a->sh_degree > 3) return fail(-1, "gft_forward: sh_degree must be 0..3");
  ViewDims dims[GFT_MAX_VIEWS];
  size_t T_total = 0, N_total = 0;
  if (!view_dims(a->views, NV, dims, T_total, N_total)) return fail(-1, "gft_forward: bad P / width / height");
  for (int i = 0; i < NV; ++i) {
    const GftViewArgs& v = a->views[i];
    if (!v.out_color || !v.out_phasor || !v.out_depth || !v.out_acc || !v.out_depth_distortion ||
        !v.out_distribution || !v.radii || !v.pixels || !v.background)
      return fail(-1, "gft_forward: required output/background pointer is null");
  }

  if (P == 0) {
    // rasterize_points.cu:104 — nothing is rendered, outputs stay at their zero fill, R = 0
    for (int i = 0; i < NV; ++i) {
      const GftViewArgs& v = a->views[i];
      const size_t N = (size_t)v.width * v.height;
      cudaMemsetAsync(v.out_color, 0, 3 * N * 4, stream);
      cudaMemsetAsync(v.out_phasor, 0, 7 * N * 4, stream);
      cudaMemsetAsync(v.out_depth, 0, N * 4, stream);
      cudaMemsetAsync(v.out_acc, 0, N * 4, stream);
      cudaMemsetAsync(v.out_depth_distortion, 0, N * 4, stream);
      cudaMemsetAsync(v.out_distribution, 0, 3 * N * 4, stream);
      if (v.out_normal) cudaMemsetAsync(v.out_normal, 0, 3 * N * 4, stream);
      if (v.out_entropy) cudaMemsetAsync(v.out_entropy, 0, N * 4, stream);
      if (v.out_amp_distortion) cudaMemsetAsync(v.out_amp_distortion, 0, N * 4, stream);
    }
    GFT_CUDA_OK("forward(P=0)");
    return 0;
  }
  if (!a->means3D || !a->opacities) return fail(-1, "gft_forward: required input pointer is null");
  for (int i = 0; i < NV; ++i)
    if (!a->views[i].viewmatrix || !a->views[i].projmatrix || !a->views[i].campos)
      return fail(-1, "gft_forward: required input pointer is null");
  if (!a->cov3D_precomp && (!a->scales || !a->rotations))
    return fail(-1, "gft_forward: need scales+rotations or cov3D_precomp");
  // rasterizer_impl.cu:269-272 (NUM_CHANNELS is 3, so this cannot fire; colours may be absent
  // only on paths that do not read them)
  if (a->shs && a->M < (a->sh_degree + 1) * (a->sh_degree + 1))
    return fail(-1, "gft_forward: shs has fewer coefficients than the active degree needs");
  if (a->shs_p && a->M_p < (a->sh_degree + 1) * (a->sh_degree + 1))
    return fail(-1, "gft_forward: shs_p has fewer coefficients than the active degree needs");

  const GeomWs gl = geom_layout(P, NV);
  const ImgWs il = img_layout(T_total, N_total);
  char* geom = geom_alloc(ctx, gl.total);
  char* img = img_alloc(ctx, il.total);
  if (!geom || !img) return fail(-3, "gft_forward: workspace callback returned null");

  uint32_t* hdr = reinterpret_cast<uint32_t*>(img + il.hdr);
  uint32_t* counts = reinterpret_cast<uint32_t*>(img + il.counts);
  uint32_t* cursors = reinterpret_cast<uint32_t*>(img + il.cursors);
  uint32_t* starts = reinterpret_cast<uint32_t*>(img + il.starts);
  uint2* ranges = reinterpret_cast<uint2*>(img + il.ranges);
  uint32_t* order = reinterpret_cast<uint32_t*>(img + il.order);
  unsigned long long* scan_state = reinterpret_cast<unsigned long long*>(img + il.hdr + 256);
  // header + scan state + sub-bin counters + cursors are adjacent: one fill
  cudaMemsetAsync(hdr, 0, il.starts - il.hdr, stream);

  gft::PreprocessParams pp;
  std::memset(&pp, 0, sizeof(pp));
  pp.P = P; pp.D = a->sh_degree; pp.M = a->M; pp.M_p = a->M_p; pp.nviews = NV;
  pp.means3D = a->means3D; pp.scales = a->scales; pp.scale_modifier = a->scale_modifier;
  pp.rotations = a->rotations; pp.opacities = a->opacities; pp.shs = a->shs; pp.shs_p = a->shs_p;
  pp.cov3D_precomp = a->cov3D_precomp; pp.colors_precomp = a->colors_precomp;
  pp.phasors_precomp = a->phasors_precomp;
  pp.prefiltered = a->prefiltered;
  pp.subtile_cull = subtile_cull_enabled();
  pp.g.cov3D = reinterpret_cast<float*>(geom + gl.cov3D);
  pp.g.rec = reinterpret_cast<float*>(geom + gl.rec);
  pp.g.depths = reinterpret_cast<float*>(geom + gl.depths);
  pp.g.tiles_touched = reinterpret_cast<uint32_t*>(geom + gl.tiles_touched);
  pp.g.rect = reinterpret_cast<uint16_t*>(geom + gl.rect);
  pp.g.clamped = reinterpret_cast<uint32_t*>(geom + gl.clamped);
  pp.g.pa = reinterpret_cast<float*>(geom + gl.pa);
  pp.tile_counts = counts;
  pp.sub_bins = il.S;
  for (int i = 0; i < NV; ++i) fill_cam(pp.views[i], a->views[i], dims[i]);
  { Stage st("preprocess_fwd", stream); gft::launch_preprocess_fwd(pp, stream); }
  GFT_CUDA_OK("preprocess");

  HostSync* hs = host_sync();
  if (!hs) return fail(-2, "gft_forward: cannot create the pinned word / event for num_rendered");

  // tile counts -> ranges + R.  R sizes the binning workspace, so it has to reach the host
  // (rasterizer_impl.cu:310-315 does a blocking cudaMemcpy): the scan kernel stores it into a
  // mapped pinned word and an event follows the kernel, so the host waits for the preprocess +
  // scan kernels only and no copy engine is involved.
  auto scan_and_post_R = [&](uint32_t cap) -> int {
    { Stage st("tile_scan", stream);
      gft::launch_tile_scan(counts, (int)T_total, il.S, cap, starts, ranges, order, hdr, scan_state,
                            hs->pinned_dev, stream); }
    GFT_CUDA_OK("tile_scan");
    const cudaError_t e = cudaEventRecord(hs->ev, stream);
    if (e != cudaSuccess)
      return fail(-2, std::string("CUDA error reading num_rendered: ") + cudaGetErrorString(e));
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
  // Everything after the scan, for a binning workspace of `cap` instances.
  auto bin_and_blend = [&](int cap) -> int {
    const BinWs bl = bin_layout(cap);
    char* bin = binning_alloc(ctx, bl.total);
    if (!bin) return fail(-3, "gft_forward: binning workspace callback returned null");
    uint32_t* point_list = reinterpret_cast<uint32_t*>(bin + bl.point_list);
    unsigned long long* entries = reinterpret_cast<unsigned long long*>(bin + bl.entries);
    if (cap > 0) {
      { Stage st("scatter_entries", stream);
        gft::launch_scatter_entries(pp, starts, cursors, entries, (uint32_t)cap, stream); }
      GFT_CUDA_OK("scatter_entries");
      { Stage st("tile_sort", stream);
        gft::launch_tile_sort(ranges, (int)T_total, entries, point_list, (int)((size_t)cap / T_total), stream); }
      GFT_CUDA_OK("tile_sort");
    }
    gft::BlendFwdParams bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.nviews = NV; bp.T_total = (int)T_total;
    bp.ranges = ranges; bp.point_list = point_list;
    bp.order = gft::option(gft::OPT_TILE_ORDER) ? order : nullptr;
    bp.img_state = reinterpret_cast<float4*>(img + il.state);
    for (int i = 0; i < NV; ++i) {
      const GftViewArgs& v = a->views[i];
      gft::BlendViewFwd& w = bp.views[i];
      w.W = v.width; w.H = v.height; w.grid_x = dims[i].gx; w.grid_y = dims[i].gy;
      w.tile_base = dims[i].tile_base; w.bg_mode = v.bg_mode; w.pix_base = dims[i].pix_base;
      w.rec = reinterpret_cast<const float4*>(pp.g.rec + (size_t)i * P * GFT_REC_FLOATS);
      w.bg = v.background;
      w.out_color = v.out_color; w.out_phasor = v.out_phasor; w.out_depth = v.out_depth;
      w.out_normal = v.out_normal; w.out_acc = v.out_acc; w.out_entropy = v.out_entropy;
      w.out_depth_distortion = v.out_depth_distortion;
      w.out_amp_distortion = v.out_amp_distortion; w.out_distribution = v.out_distribution;
      w.pixels = v.pixels;
    }
    { Stage st("blend_fwd", stream); gft::launch_blend_fwd(bp, stream); }
    GFT_CUDA_OK("blend_fwd");
    return 0;
  };

  int R = 0;
  if (a->R_hint <= 0) {
    // exact mode (reference behaviour): wait for the count, then size the workspace from it
    int rc = scan_and_post_R(0xffffffffu);
    if (rc < 0) return rc;
    rc = wait_R(R);
    if (rc < 0) return rc;
    rc = bin_and_blend(R);
    if (rc < 0) return rc;
  } else {
    // hinted mode: the caller's estimate sizes the workspace, all kernels are enqueued at once
    // (ranges clamped to the estimate on the device); the host only learns R afterwards — no
    // bubble in the stream.
    int rc = scan_and_post_R((uint32_t)a->R_hint);
    if (rc < 0) return rc;
    rc = bin_and_blend(a->R_hint);
    if (rc < 0) return rc;
    rc = wait_R(R);
    if (rc < 0) return rc;
    if (R > a->R_hint) {   // estimate too small: redo scan + binning + blend with the exact size
      for (int i = 0; i < NV; ++i)
        cudaMemsetAsync(a->views[i].pixels, 0, (size_t)P * sizeof(float), stream);
      cudaMemsetAsync(hdr, 0, il.counts - il.hdr, stream);                 // ticket + scan state
      cudaMemsetAsync(cursors, 0, T_total * (size_t)il.S * 4, stream);
      rc = scan_and_post_R(0xffffffffu);
      if (rc < 0) return rc;
      rc = bin_and_blend(R);
      if (rc < 0) return rc;
    }
  }
  return R;
}

int gft_backward_views(const GftBackwardViewsArgs* a, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(-1, "gft_backward: null args");
  const bool debug = a->debug != 0;
  const int P = a->P, NV = a->n_views;
  if (P < 0) return fail(-1, "gft_backward: bad P");
  if (NV < 1 || NV > GFT_MAX_VIEWS || !a->views)
    return fail(-1, "gft_backward: n_views must be 1.." + std::to_string(GFT_MAX_VIEWS));
  ViewDims dims[GFT_MAX_VIEWS];
  size_t T_total = 0, N_total = 0;
  if (!view_dims(a->views, NV, dims, T_total, N_total)) return fail(-1, "gft_backward: bad P / width / height");
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
  if (!a->geom_buffer || !a->img_buffer || !a->means3D)
    return fail(-1, "gft_backward: required input pointer is null");
  if (a->R > 0 && !a->binning_buffer) return fail(-1, "gft_backward: binning buffer is null");
  for (int i = 0; i < NV; ++i) {
    const GftViewArgs& v = a->views[i];
    if (!v.radii || !v.viewmatrix || !v.projmatrix || !v.campos || !v.background)
      return fail(-1, "gft_backward: required input pointer is null");
    if (!v.dL_dout_color || !v.dL_dout_phasor || !v.dL_dout_depth || !v.dL_dout_acc ||
        !v.dL_dout_depth_distortion)
      return fail(-1, "gft_backward: incoming gradient pointer is null");
    if (!v.dL_dmeans2D) return fail(-1, "gft_backward: required output pointer is null");
  }
  if (!a->dL_dopacity || !a->dL_dmeans3D || !a->scratch)
    return fail(-1, "gft_backward: required output pointer is null");
  if ((a->shs && !a->dL_dsh) || (a->shs_p && !a->dL_dsh_p) ||
      (a->scales && (!a->dL_dscales || !a->dL_drotations)))
    return fail(-1, "gft_backward: missing gradient buffer for a provided input");

  const GeomWs gl = geom_layout(P, NV);
  const ImgWs il = img_layout(T_total, N_total);
  const char* geom = a->geom_buffer;
  const char* img = a->img_buffer;
  const char* bin = a->binning_buffer;

  { Stage st("zero_grad_records", stream);
    cudaMemsetAsync(a->scratch, 0, (size_t)P * NV * GFT_GRAD_FLOATS * 4, stream); }

  if (a->R > 0) {
    gft::BlendBwdParams bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.nviews = NV; bp.T_total = (int)T_total;
    bp.ranges = reinterpret_cast<const uint2*>(img + il.ranges);
    bp.order = gft::option(gft::OPT_TILE_ORDER) ? reinterpret_cast<const uint32_t*>(img + il.order) : nullptr;
    bp.point_list = reinterpret_cast<const uint32_t*>(bin);  // sorted list: offset 0 of the workspace
    bp.img_state = reinterpret_cast<const float4*>(img + il.state);
    for (int i = 0; i < NV; ++i) {
      const GftViewArgs& v = a->views[i];
      gft::BlendViewBwd& w = bp.views[i];
      w.W = v.width; w.H = v.height; w.grid_x = dims[i].gx; w.grid_y = dims[i].gy;
      w.tile_base = dims[i].tile_base; w.bg_mode = v.bg_mode; w.pix_base = dims[i].pix_base;
      w.rec = reinterpret_cast<const float4*>(geom + gl.rec) + (size_t)i * P * (GFT_REC_FLOATS / 4);
      w.bg = v.background;
      w.dL_dcolor = v.dL_dout_color; w.dL_dphasor = v.dL_dout_phasor;
      w.dL_ddepth = v.dL_dout_depth; w.dL_dacc = v.dL_dout_acc; w.dL_ddd = v.dL_dout_depth_distortion;
      w.grad_rec = a->scratch + (size_t)i * P * GFT_GRAD_FLOATS;
    }
    { Stage st("blend_bwd", stream); gft::launch_blend_bwd(bp, stream); }
    GFT_CUDA_OK("blend_bwd");
  }

  gft::PreprocessBwdParams pb;
  std::memset(&pb, 0, sizeof(pb));
  pb.P = P; pb.D = a->sh_degree; pb.M = a->M; pb.M_p = a->M_p; pb.nviews = NV;
  pb.means3D = a->means3D; pb.shs = a->shs; pb.shs_p = a->shs_p;
  pb.scales = a->scales; pb.rotations = a->rotations; pb.scale_modifier = a->scale_modifier;
  // rasterizer_impl.cu:471
  pb.cov3D = a->cov3D_precomp ? a->cov3D_precomp : reinterpret_cast<const float*>(geom + gl.cov3D);
  pb.rec = reinterpret_cast<const float*>(geom + gl.rec);
  pb.clamped = reinterpret_cast<const uint32_t*>(geom + gl.clamped);
  pb.pa = reinterpret_cast<const float*>(geom + gl.pa);
  pb.grad_rec = a->scratch;
  pb.accumulate = a->accumulate;
  pb.dL_dopacity = a->dL_dopacity;
  pb.dL_dmeans3D = a->dL_dmeans3D; pb.dL_dsh = a->dL_dsh; pb.dL_dsh_p = a->dL_dsh_p;
  pb.dL_dscales = a->dL_dscales; pb.dL_drotations = a->dL_drotations;
  pb.dL_dphase_offset = a->dL_dphase_offset; pb.dL_ddc_offset = a->dL_ddc_offset;
  pb.dL_dcolors = a->dL_dcolors; pb.dL_dcov3D = a->dL_dcov3D;
  pb.dL_dconic = a->dL_dconic; pb.dL_ddist = a->dL_ddist; pb.dL_dndc = a->dL_dndc;
  for (int i = 0; i < NV; ++i) fill_cam(pb.views[i], a->views[i], dims[i]);
  { Stage st("preprocess_bwd", stream); gft::launch_preprocess_bwd(pb, stream); }
  GFT_CUDA_OK("preprocess_bwd");
  return 0;
}

// ---- the reference's single-view call shape: a batch of one ---------------------------------------
int gft_forward(const GftForwardArgs* a, gft_alloc_fn geom_alloc, gft_alloc_fn binning_alloc,
                gft_alloc_fn img_alloc, void* ctx, gft_stream_t stream) {
  if (!a) return fail(-1, "gft_forward: null args");
  if (a->P < 0 || a->width <= 0 || a->height <= 0) return fail(-1, "gft_forward: bad P / width / height");
  GftViewArgs v;
  std::memset(&v, 0, sizeof(v));
  v.width = a->width; v.height = a->height;
  v.background = a->background; v.bg_mode = a->bg_mode;
  v.viewmatrix = a->viewmatrix; v.projmatrix = a->projmatrix; v.campos = a->campos;
  v.tan_fovx = a->tan_fovx; v.tan_fovy = a->tan_fovy;
  v.near_n = a->near_n; v.far_n = a->far_n; v.depth_range = a->depth_range;
  v.use_view_dependent_phase = a->use_view_dependent_phase;
  v.phase_offset = a->phase_offset; v.dc_offset = a->dc_offset;
  v.out_color = a->out_color; v.out_phasor = a->out_phasor; v.out_depth = a->out_depth;
  v.out_normal = a->out_normal; v.out_acc = a->out_acc; v.out_entropy = a->out_entropy;
  v.out_depth_distortion = a->out_depth_distortion; v.out_amp_distortion = a->out_amp_distortion;
  v.pixels = a->pixels; v.out_distribution = a->out_distribution; v.radii = a->radii;
  GftForwardViewsArgs b;
  std::memset(&b, 0, sizeof(b));
  b.P = a->P; b.sh_degree = a->sh_degree; b.M = a->M; b.M_p = a->M_p; b.n_views = 1;
  b.means3D = a->means3D; b.shs = a->shs; b.shs_p = a->shs_p;
  b.colors_precomp = a->colors_precomp; b.phasors_precomp = a->phasors_precomp;
  b.opacities = a->opacities; b.scales = a->scales; b.scale_modifier = a->scale_modifier;
  b.rotations = a->rotations; b.cov3D_precomp = a->cov3D_precomp;
  b.prefiltered = a->prefiltered; b.debug = a->debug;
  b.views = &v; b.R_hint = a->R_hint;
  return gft_forward_views(&b, geom_alloc, binning_alloc, img_alloc, ctx, stream);
}

int gft_backward(const GftBackwardArgs* a, gft_stream_t stream) {
  if (!a) return fail(-1, "gft_backward: null args");
  if (a->P < 0 || a->width <= 0 || a->height <= 0) return fail(-1, "gft_backward: bad P / width / height");
  if (a->dL_dphasors)
    return fail(-1, "gft_backward: dL_dphasors is a reference-only field (the 7 phasor gradients are "
                    "reduced to the 4 combinations the phasor backward consumes); pass NULL");
  GftViewArgs v;
  std::memset(&v, 0, sizeof(v));
  v.width = a->width; v.height = a->height;
  v.background = a->background; v.bg_mode = a->bg_mode;
  v.viewmatrix = a->viewmatrix; v.projmatrix = a->projmatrix; v.campos = a->campos;
  v.tan_fovx = a->tan_fovx; v.tan_fovy = a->tan_fovy;
  v.near_n = a->near_n; v.far_n = a->far_n; v.depth_range = a->depth_range;
  v.use_view_dependent_phase = a->use_view_dependent_phase;
  v.phase_offset = a->phase_offset; v.dc_offset = a->dc_offset;
  v.radii = const_cast<int*>(a->radii);
  v.dL_dout_color = a->dL_dout_color; v.dL_dout_phasor = a->dL_dout_phasor;
  v.dL_dout_depth = a->dL_dout_depth; v.dL_dout_acc = a->dL_dout_acc;
  v.dL_dout_depth_distortion = a->dL_dout_depth_distortion;
  v.dL_dmeans2D = a->dL_dmeans2D;
  GftBackwardViewsArgs b;
  std::memset(&b, 0, sizeof(b));
  b.P = a->P; b.sh_degree = a->sh_degree; b.M = a->M; b.M_p = a->M_p; b.R = a->R; b.n_views = 1;
  b.means3D = a->means3D; b.shs = a->shs; b.shs_p = a->shs_p;
  b.colors_precomp = a->colors_precomp; b.phasors_precomp = a->phasors_precomp;
  b.scales = a->scales; b.scale_modifier = a->scale_modifier; b.rotations = a->rotations;
  b.cov3D_precomp = a->cov3D_precomp;
  b.geom_buffer = a->geom_buffer; b.binning_buffer = a->binning_buffer; b.img_buffer = a->img_buffer;
  b.views = &v;
  b.dL_dopacity = a->dL_dopacity; b.dL_dmeans3D = a->dL_dmeans3D;
  b.dL_dsh = a->dL_dsh; b.dL_dsh_p = a->dL_dsh_p;
  b.dL_dscales = a->dL_dscales; b.dL_drotations = a->dL_drotations;
  b.dL_dphase_offset = a->dL_dphase_offset; b.dL_ddc_offset = a->dL_ddc_offset;
  b.dL_dcolors = a->dL_dcolors; b.dL_dcov3D = a->dL_dcov3D;
  b.dL_dconic = a->dL_dconic; b.dL_ddist = a->dL_ddist; b.dL_dndc = a->dL_dndc;
  b.scratch = a->scratch; b.debug = a->debug; b.accumulate = a->accumulate;
  return gft_backward_views(&b, stream);
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
