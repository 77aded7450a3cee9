// kernels.h — internal launch interfaces between api.cu and the kernel translation units.
//
// Every kernel of the rasterizer works on a BATCH of views of the same Gaussian set (the colour
// and the ToF camera of an iteration, gaussian_renderer/__init__.py:107-128; the 8 cameras of a
// multi-view batch; the frames of a render-only sweep): one launch covers all views, the tiles of
// all views form one global tile index space (view v owns tiles [tile_base_v, tile_base_v + T_v)),
// and the per-Gaussian parameters are read once per launch instead of once per view.  A single
// view (the reference's call shape) is a batch of one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define GFT_MAX_VIEWS 16

namespace gft {

// process-wide count of kernels launched by this library (gft_launch_count in the C ABI)
void note_launches(int n);

// Opt a kernel in to more than 48 KB of dynamic shared memory, once per device (the attribute is
// per device; setting it on every launch costs a driver call each time).
template <typename K>
inline void ensure_dynamic_smem(K kernel, int bytes, unsigned long long* done_mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
}

// Number of SMs of the current device (cached per device).
int sm_count();

// Tunables / A-B switches (gft_set_option in the C ABI; defaults from the environment, read once).
enum { OPT_SORT_CAP = 0, OPT_BWD_PRED, OPT_PBWD_MINB, OPT_NO_CULL, OPT_SORT_RADIX, OPT_SUB_BINS, OPT_SORT_MATCH, OPT_TILE_ORDER, OPT_BWD_RING, OPT_PFWD_MINB, OPT_BLEND_HALF, OPT_SORT_ADAPT, OPT_COUNT };
int option(int id);

// Records the thread-local error string behind gft_last_error() and returns `code` (api.cu).
int set_error(int code, const char* msg);

// Warps per blend block (8 = a whole 16x16 tile, 4 = a 16x8 half); see blend_fwd.cu.
// GFT_BLEND_WARPS overrides the choice (for A/B measurements).
int blend_block_warps(int tiles);

// ---- per-view constants ------------------------------------------------------------------------
struct ViewCam {
  int W, H, grid_x, grid_y;
  int tile_base;                 // first global tile of this view
  int use_view_dependent_phase;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  float near_n, far_n, dist2phase, phase_offset, dc_offset;
  int pad_;
  const float* viewmatrix;       // 16 floats, column-major
  const float* projmatrix;       // 16 floats, full view*proj
  const float* campos;           // 3 floats
  int* radii;                    // [P] (forward: written; backward: read)
  float* pixels;                 // [P] forward only
  float* dL_dmeans2D;            // [P,3] backward only
};

// Per-view slices of the geometry workspace are laid out view-major: array[v][P][...].
struct GeomViews {
  float* cov3D;            // [P][6]            shared by all views (rasterizer_impl.cu:471)
  float* rec;              // [V][P][20]        blend records
  float* depths;           // [V][P]            view-space z (the sort key's depth bits)
  uint32_t* tiles_touched; // [V][P]
  uint16_t* rect;          // [V][P][4]
  uint32_t* clamped;       // [V][P]            packed r | g<<8 | b<<16 | amp<<24
  float* pa;               // [V][P][2]
};

struct PreprocessParams {
  int P, D, M, M_p, nviews;
  const float* means3D;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* opacities;
  const float* shs;
  const float* shs_p;
  const float* cov3D_precomp;
  const float* colors_precomp;
  const float* phasors_precomp;
  int prefiltered;
  int subtile_cull;        // 0: write infinite extents (blend kernels then test every pair)
  GeomViews g;
  uint32_t* tile_counts;   // [T_total][sub_bins] zero-filled before the launch; += 1 per (Gaussian, tile) instance
  int sub_bins;            // sub-counters per tile (power of two): Gaussian g counts into sub-bin g % sub_bins
  ViewCam views[GFT_MAX_VIEWS];
};

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* V, uint8_t* present,
                         float near_n, float far_n, cudaStream_t stream);

// ---- binning: tile-segmented (binning.cu) ------------------------------------------------------
// tile_counts [T_total][S] -> starts [T_total*S + 1] (exclusive scan, chained over blocks through
// `scan_state`; hdr[0] is its ticket counter), ranges (empty tiles read (0,0) like the reference's
// memset + identifyTileRanges, rasterizer_impl.cu:118-140,341), the instance count R in hdr[1] and
// `order`, the global tile ids by descending list length (launch order of the blend kernels).
// hdr, scan_state and the cursors must be zero on entry.  Everything is clamped to `capacity`
// (only ever effective when a caller's size hint was too small; the forward is then repeated).
void launch_tile_scan(const uint32_t* tile_counts, int T_total, int sub_bins, uint32_t capacity,
                      uint32_t* starts, uint2* ranges, uint32_t* order, uint32_t* hdr,
                      unsigned long long* scan_state, uint32_t* host_R, cudaStream_t stream);
// Every (view, Gaussian, tile) instance writes its entry (float_bits(view_z) << 32 | Gaussian id)
// into its tile's segment, at a slot handed out by its sub-bin's atomic cursor (any order).
// `capacity` is the clamp the scan was run with (= the number of entries the buffer holds).
void launch_scatter_entries(const PreprocessParams& pp, const uint32_t* starts, uint32_t* cursors,
                            unsigned long long* entries, uint32_t capacity, cudaStream_t stream);
// Sorts every tile segment on the 64-bit entry (depth bits, then Gaussian id): exactly the order
// of the reference's stable radix sort of (tile << 32 | depth) keys over instances emitted in
// ascending Gaussian order (rasterizer_impl.cu:90-110,334-339).  Writes the sorted entries back
// and the Gaussian ids to point_list.
void launch_tile_sort(const uint2* ranges, int T_total, unsigned long long* entries,
                      uint32_t* point_list, int mean_len_hint, cudaStream_t stream);

// ---- radix sort (knn + the A/B hooks) ----------------------------------------------------------
size_t sort_pairs_temp_bytes(int R);
int sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
               const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
               cudaStream_t stream, const uint32_t* d_R = nullptr);
bool sort_result_in_out(int end_bit);

// ---- blend -----------------------------------------------------------------------------------
struct BlendViewFwd {
  int W, H, grid_x, grid_y;
  int tile_base, bg_mode;
  size_t pix_base;          // first pixel of this view in img_state
  const float4* rec;        // this view's blend records
  const float* bg;
  float* out_color;
  float* out_phasor;
  float* out_depth;
  float* out_normal;
  float* out_acc;
  float* out_entropy;
  float* out_depth_distortion;
  float* out_amp_distortion;
  float* out_distribution;
  float* pixels;
};
struct BlendFwdParams {
  int nviews, T_total;
  const uint32_t* order;   // block b works on global tile order[b] (nullptr: b)
  const uint2* ranges;
  const uint32_t* point_list;
  float4* img_state;  // final_T, w_z_total, w_z2_total, n_contrib bits
  BlendViewFwd views[GFT_MAX_VIEWS];
};
void launch_blend_fwd(const BlendFwdParams& p, cudaStream_t stream);

struct BlendViewBwd {
  int W, H, grid_x, grid_y;
  int tile_base, bg_mode;
  size_t pix_base;
  const float4* rec;
  const float* bg;
  const float* dL_dcolor;
  const float* dL_dphasor;
  const float* dL_ddepth;
  const float* dL_dacc;
  const float* dL_ddd;
  float* grad_rec;    // [P][16] of this view, zero-filled before launch
};
struct BlendBwdParams {
  int nviews, T_total;
  const uint32_t* order;
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* img_state;
  BlendViewBwd views[GFT_MAX_VIEWS];
};
void launch_blend_bwd(const BlendBwdParams& p, cudaStream_t stream);

// ---- preprocess backward ---------------------------------------------------------------------
struct PreprocessBwdParams {
  int P, D, M, M_p, nviews;
  const float* means3D;
  const float* shs;
  const float* shs_p;
  const float* scales;
  const float* rotations;
  float scale_modifier;
  const float* cov3D;      // precomp or saved, [P][6]
  const float* rec;        // [V][P][20] forward blend records (conic, dist)
  const uint32_t* clamped; // [V][P]
  const float* pa;         // [V][P][2]
  const float* grad_rec;   // [V][P][16] from blend backward
  int accumulate;          // 0 overwrite, 1 add, 2 atomic add (see include/gftorf.h)
  // outputs: the parameter gradients, summed over the views of the batch, each row written once
  float* dL_dopacity;
  float* dL_dmeans3D;
  float* dL_dsh;
  float* dL_dsh_p;
  float* dL_dscales;
  float* dL_drotations;
  float* dL_dphase_offset;
  float* dL_ddc_offset;
  float* dL_dcolors;       // optional: grad of colors_precomp (summed over views)
  float* dL_dcov3D;        // optional: grad of cov3D_precomp (summed over views)
  float* dL_dconic;        // optional per-view intermediates of view 0 (single-view debugging)
  float* dL_ddist;
  float* dL_dndc;
  ViewCam views[GFT_MAX_VIEWS];
};
void launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream);

// ---- knn -------------------------------------------------------------------------------------
size_t knn_workspace_bytes(int P);
int knn_dist2(const float* points, int P, float* out, char* workspace, cudaStream_t stream);

}  // namespace gft
