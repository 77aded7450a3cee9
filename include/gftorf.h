/*
 * gftorf.h — C ABI of the B200-native RGB+ToF Gaussian rasterizer and 3-NN initialiser.
 *
 * This is the drop-in boundary for the two CUDA submodules of brownvc/gftorf.  Every entry point
 * below replaces one function of the reference's pybind11 modules; the reference interface each
 * one stands in for is cited as file:line relative to the reference checkout:
 *
 *   rasterizer = submodules/diff-gaussian-rasterization-w-tof
 *   knn        = submodules/simple-knn
 *
 * Rules of the boundary
 *   - plain C: POD structs of raw device pointers, ints and floats; no torch, no pybind, no C++.
 *   - the library owns no memory across calls.  Outputs are caller-allocated; the three opaque
 *     workspaces (geometry / binning / image) are obtained through caller-supplied callbacks
 *     because the binning size is only known after the tile count scan — exactly the shape of the
 *     reference (rasterizer/cuda_rasterizer/rasterizer.h:34-36, rasterize_points.cu:27-33).
 *   - "absent" optional inputs are NULL pointers (reference: empty tensors whose data_ptr is
 *     null, diff_gaussian_rasterization_w_tof/__init__.py:239-254).
 *   - every kernel is launched on the caller's stream; the library is re-entrant (no globals
 *     except a thread-local error string).
 *   - return value < 0 means failure; gft_last_error() gives the message.
 */
#ifndef GFTORF_H_INCLUDED
#define GFTORF_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFT_ABI_VERSION 2

/* Channel counts, rasterizer/cuda_rasterizer/config.h:15-23 */
#define GFT_NUM_CHANNELS 3         /* RGB */
#define GFT_NUM_CHANNELS_CWTOF 2   /* phase, amplitude */
#define GFT_NUM_CHANNELS_PHASOR 7  /* real, imag, amp, quad0..quad3 */
#define GFT_TILE 16                /* BLOCK_X == BLOCK_Y */

/* Workspace callback: return a device pointer to `bytes` bytes (>=256-B aligned), valid until the
 * caller frees it.  Mirrors std::function<char*(size_t)> of rasterizer.h:34-36. */
typedef char* (*gft_alloc_fn)(void* ctx, size_t bytes);

/* CUDA stream handle (cudaStream_t) passed as an opaque pointer so this header needs no CUDA. */
typedef void* gft_stream_t;

/* ------------------------------------------------------------------------------------------
 * Forward.  Replaces RasterizeGaussiansCUDA (rasterize_points.cu:35-165) +
 * CudaRasterizer::Rasterizer::forward (cuda_rasterizer/rasterizer_impl.cu:215-378).
 * ---------------------------------------------------------------------------------------- */
typedef struct GftForwardArgs {
  int P;          /* number of Gaussians */
  int sh_degree;  /* active SH degree D (0..3) */
  int M;          /* SH coefficients per Gaussian in `shs`   (0 if absent) */
  int M_p;        /* SH coefficients per Gaussian in `shs_p` (0 if absent) */
  int width, height;

  /* Background.  bg_mode 0: per-pixel map, read as background[ch*H*W + pix] for ch 0..6 with the
   * RENDER H,W as plane stride (forward.cu:644,649; quirk A.7-3 of SURVEY.md).
   * bg_mode 1: background holds 7 per-channel constants (a stride-0 expanded map, train.py:127)
   * — same values, no materialisation. */
  const float* background;
  int bg_mode;

  const float* means3D;         /* [P,3] */
  const float* shs;             /* [P,M,3]   or NULL */
  const float* shs_p;           /* [P,M_p,2] or NULL */
  const float* colors_precomp;  /* [P,3]     or NULL */
  const float* phasors_precomp; /* [P,2]     or NULL */
  const float* opacities;       /* [P] */
  const float* scales;          /* [P,3]     or NULL */
  float scale_modifier;
  const float* rotations;       /* [P,4]     or NULL */
  const float* cov3D_precomp;   /* [P,6]     or NULL */
  const float* viewmatrix;      /* 16 floats on device, column-major (auxiliary.h:61-80) */
  const float* projmatrix;      /* 16 floats on device, full view*proj */
  const float* campos;          /* 3 floats on device */
  float tan_fovx, tan_fovy;
  int prefiltered;
  int debug;                    /* sync + check after every stage (auxiliary.h:215-222) */
  float near_n, far_n, depth_range;
  int use_view_dependent_phase;
  float phase_offset, dc_offset;

  /* Outputs, all caller-allocated, need NOT be pre-zeroed: every element is written. */
  float* out_color;            /* [3,H,W] */
  float* out_phasor;           /* [7,H,W] */
  float* out_depth;            /* [1,H,W] */
  float* out_normal;           /* [3,H,W] always 0 (forward.cu:667 commented)   — may be NULL */
  float* out_acc;              /* [1,H,W] */
  float* out_entropy;          /* [1,H,W] always 0 (forward.cu:655-658)         — may be NULL */
  float* out_depth_distortion; /* [1,H,W] */
  float* out_amp_distortion;   /* [1,H,W] always 0 (forward.cu:665)             — may be NULL */
  float* pixels;               /* [P]  number of pixels each Gaussian contributed to */
  float* out_distribution;     /* [3,H,W] first hit: alpha, dist, amp */
  int* radii;                  /* [P] */

  /* 0: exact mode (reference behaviour) — the host waits for the instance count R before it sizes
   * the binning workspace.  > 0: an estimate of R (e.g. 1.25x the previous call's): the workspace
   * is sized from it, every kernel is enqueued at once and reads R on the device, and the host
   * learns R afterwards; if R exceeds the estimate the binning callback is called a second time
   * with the exact size and the tail of the pipeline re-runs.  Results are identical. */
  int R_hint;
} GftForwardArgs;

/* Returns num_rendered R (>=0), or <0 on error.  The three callbacks are called exactly once
 * each (geom, img before the first kernel; binning after the tile-count scan — a second time
 * only when R_hint was too small).  A batch of one view: see gft_forward_views. */
int gft_forward(const GftForwardArgs* args,
                gft_alloc_fn geom_alloc, gft_alloc_fn binning_alloc, gft_alloc_fn img_alloc,
                void* alloc_ctx, gft_stream_t stream);

/* Workspace sizes, for callers that pre-allocate (render-only path).  Mirrors
 * CudaRasterizer::required<T>() (cuda_rasterizer/rasterizer_impl.h:76-82). */
size_t gft_geom_bytes(int P);
size_t gft_img_bytes(int width, int height);
size_t gft_binning_bytes(int R);
size_t gft_geom_bytes_views(int P, int n_views);
size_t gft_img_bytes_views(int n_views, const int* widths, const int* heights);

/* ------------------------------------------------------------------------------------------
 * Backward.  Replaces RasterizeGaussiansBackwardCUDA (rasterize_points.cu:167-281) +
 * CudaRasterizer::Rasterizer::backward (cuda_rasterizer/rasterizer_impl.cu:382-499).
 * ---------------------------------------------------------------------------------------- */
typedef struct GftBackwardArgs {
  int P, sh_degree, M, M_p, R;
  int width, height;
  const float* background;
  int bg_mode;
  const float* means3D;
  const float* shs;
  const float* shs_p;
  const float* colors_precomp;
  const float* phasors_precomp;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* cov3D_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float tan_fovx, tan_fovy;
  const int* radii;
  const char* geom_buffer;     /* the three workspaces handed out by gft_forward */
  const char* binning_buffer;
  const char* img_buffer;

  /* Incoming pixel gradients.  normal / entropy / amp_distortion / pixels / distribution grads
   * are ignored by the reference (backward.cu never reads them) and are not part
   * of this struct. */
  const float* dL_dout_color;            /* [3,H,W] */
  const float* dL_dout_phasor;           /* [7,H,W] */
  const float* dL_dout_depth;            /* [1,H,W] */
  const float* dL_dout_acc;              /* [1,H,W] */
  const float* dL_dout_depth_distortion; /* [1,H,W] */

  /* Outputs, caller-allocated, need NOT be pre-zeroed (rows of culled Gaussians are written 0).
   * Required: */
  float* dL_dmeans2D;       /* [P,3] (z = 0) */
  float* dL_dopacity;       /* [P] */
  float* dL_dmeans3D;       /* [P,3] */
  float* dL_dsh;            /* [P,M,3]    or NULL when shs absent */
  float* dL_dsh_p;          /* [P,M_p,2]  or NULL when shs_p absent */
  float* dL_dscales;        /* [P,3]      or NULL when scales absent */
  float* dL_drotations;     /* [P,4]      or NULL when scales absent */
  float* dL_dphase_offset;  /* [1] */
  float* dL_ddc_offset;     /* [1] */
  /* Optional intermediates (reference materialises them always, rasterize_points.cu:222-236);
   * NULL = keep in registers only: */
  float* dL_dcolors;        /* [P,3]  grad of colors_precomp */
  float* dL_dphasors;       /* [P,7]  REFERENCE-ONLY: must be NULL for gft_backward (the blend
                               backward reduces the 7 phasor gradients to the 4 combinations the
                               phasor backward consumes); used by the oracle shim */
  float* dL_dcov3D;         /* [P,6]  grad of cov3D_precomp */
  float* dL_dconic;         /* [P,4]  (x, y, unused, w) as the reference's [P,2,2] */
  float* dL_ddist;          /* [P] */
  float* dL_dndc;           /* [P] */

  /* Scratch for the per-Gaussian blend-gradient records, >= gft_backward_scratch_bytes(P) bytes.
   * Zero-filled by the library. */
  float* scratch;

  int debug;
  float near_n, far_n, depth_range;
  int use_view_dependent_phase;
  float phase_offset, dc_offset;

  /* 0 (reference semantics): every output is overwritten.  1: dL_dmeans3D, dL_dsh, dL_dsh_p,
   * dL_dopacity, dL_dscales, dL_drotations, dL_dphase_offset and dL_ddc_offset are ADDED TO
   * (rows of culled Gaussians untouched), so the views of a multi-camera batch accumulate
   * straight into one gradient bucket (the reference gets the same sum from autograd's
   * AccumulateGrad, train.py:279).  dL_dmeans2D and the optional intermediates stay per view.
   * 2: as 1 with ATOMIC adds, so backward calls of different views may run concurrently (on
   * different streams) into the same bucket; the bucket must have been zero-filled. */
  int accumulate;
} GftBackwardArgs;

size_t gft_backward_scratch_bytes(int P);
int gft_backward(const GftBackwardArgs* args, gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Batched views.  One call rasterizes the SAME Gaussians for up to GFT_MAX_VIEWS cameras — the
 * colour and the ToF camera of a training iteration (gaussian_renderer/__init__.py:107-128 makes
 * two rasterizer calls back to back), the cameras of a multi-view batch, the frames of a
 * render-only sweep (render.py:95-209).  Every kernel covers all views in one launch, the
 * 364 B/Gaussian of parameters are read once per pass instead of once per view, and the backward
 * writes each parameter-gradient row once: the SUM over the views of the batch (what autograd's
 * AccumulateGrad produces for the reference's separate calls, train.py:279).  Per view the
 * results are exactly those of gft_forward / gft_backward, which are batches of one.
 * ---------------------------------------------------------------------------------------- */
#define GFT_MAX_VIEWS 16

typedef struct GftViewArgs {
  int width, height;
  const float* background;      /* see GftForwardArgs.background */
  int bg_mode;
  const float* viewmatrix;      /* 16 floats on device, column-major */
  const float* projmatrix;      /* 16 floats on device, full view*proj */
  const float* campos;          /* 3 floats on device */
  float tan_fovx, tan_fovy;
  float near_n, far_n, depth_range;
  int use_view_dependent_phase;
  float phase_offset, dc_offset;
  /* forward outputs of this view (as in GftForwardArgs) */
  float* out_color;
  float* out_phasor;
  float* out_depth;
  float* out_normal;            /* may be NULL */
  float* out_acc;
  float* out_entropy;           /* may be NULL */
  float* out_depth_distortion;
  float* out_amp_distortion;    /* may be NULL */
  float* pixels;                /* [P] */
  float* out_distribution;
  int* radii;                   /* [P]; written by the forward, read by the backward */
  /* backward: incoming pixel gradients of this view and its screen-space gradient output */
  const float* dL_dout_color;
  const float* dL_dout_phasor;
  const float* dL_dout_depth;
  const float* dL_dout_acc;
  const float* dL_dout_depth_distortion;
  float* dL_dmeans2D;           /* [P,3] (z = 0), per view: its norm feeds densification */
} GftViewArgs;

typedef struct GftForwardViewsArgs {
  int P, sh_degree, M, M_p, n_views;
  const float* means3D;
  const float* shs;
  const float* shs_p;
  const float* colors_precomp;
  const float* phasors_precomp;
  const float* opacities;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* cov3D_precomp;
  int prefiltered;
  int debug;
  const GftViewArgs* views;     /* host array of n_views entries */
  int R_hint;                   /* as GftForwardArgs.R_hint, for the instance count of the whole batch */
} GftForwardViewsArgs;

/* Returns the number of (Gaussian, tile) instances of the whole batch, or <0 on error. */
int gft_forward_views(const GftForwardViewsArgs* args,
                      gft_alloc_fn geom_alloc, gft_alloc_fn binning_alloc, gft_alloc_fn img_alloc,
                      void* alloc_ctx, gft_stream_t stream);

typedef struct GftBackwardViewsArgs {
  int P, sh_degree, M, M_p, R, n_views;
  const float* means3D;
  const float* shs;
  const float* shs_p;
  const float* colors_precomp;
  const float* phasors_precomp;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* cov3D_precomp;
  const char* geom_buffer;      /* the three workspaces handed out by gft_forward_views */
  const char* binning_buffer;
  const char* img_buffer;
  const GftViewArgs* views;     /* the same views, in the same order, with the dL_* fields set */
  /* parameter gradients, summed over the views (see GftBackwardArgs for shapes and `accumulate`) */
  float* dL_dopacity;
  float* dL_dmeans3D;
  float* dL_dsh;
  float* dL_dsh_p;
  float* dL_dscales;
  float* dL_drotations;
  float* dL_dphase_offset;
  float* dL_ddc_offset;
  float* dL_dcolors;            /* optional */
  float* dL_dcov3D;             /* optional */
  float* dL_dconic;             /* optional, view 0 only */
  float* dL_ddist;              /* optional, view 0 only */
  float* dL_dndc;               /* optional, view 0 only */
  float* scratch;               /* >= gft_backward_scratch_bytes_views(P, n_views) bytes */
  int debug;
  int accumulate;
} GftBackwardViewsArgs;

size_t gft_backward_scratch_bytes_views(int P, int n_views);
int gft_backward_views(const GftBackwardViewsArgs* args, gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * markVisible (rasterize_points.cu:283-304, rasterizer_impl.cu:54-68,143-159).
 * present[i] = near <= view_z(i) <= far   (x/y are NOT tested: auxiliary.h:169)
 * ---------------------------------------------------------------------------------------- */
int gft_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, float near_n, float far_n, gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * distCUDA2 (knn/spatial.cu:15-26 -> SimpleKNN::knn, knn/simple_knn.cu:185-221).
 * out[i] = mean of the squared distances to the 3 nearest other points.
 * `workspace` >= gft_dist2_workspace_bytes(P) bytes of device memory.
 * ---------------------------------------------------------------------------------------- */
size_t gft_dist2_workspace_bytes(int P);
int gft_dist2(const float* points, int P, float* out, char* workspace, gft_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Debug / test accessors: byte offsets of the arrays inside the opaque workspaces, so the
 * bit-exact tests can compare tiles_touched, sorted keys, point_list, ranges and n_contrib with
 * the reference's.
 * ---------------------------------------------------------------------------------------- */
typedef struct GftWorkspaceLayout {
  /* geometry workspace: cov3D is per Gaussian; the other arrays are view-major, [n_views][P][...]
   * (the slice of view v starts v * P * element_size bytes after the offset below) */
  size_t geom_cov3D;          /* float[P][6] */
  size_t geom_rec;            /* float[V][P][20]: x y ex ey | conA conB conC opac | r g b dist | ph0..ph3 | ph4 ph5 ph6 ndc */
  size_t geom_depths;         /* float[V][P]  view-space z */
  size_t geom_tiles_touched;  /* uint32[V][P] */
  size_t geom_rect;           /* uint16[V][P][4]: xmin ymin xmax ymax (tile units) */
  size_t geom_clamped;        /* uint8[V][P][4]: r g b amp */
  size_t geom_pa;             /* float[V][P][2]: phase_sh, amplitude */
  size_t geom_total;
  /* binning workspace (per instance) */
  size_t bin_point_list;      /* uint32[R] sorted Gaussian ids */
  size_t bin_entries;         /* uint64[R] sorted entries: float_bits(view_z) << 32 | Gaussian id; the
                                 reference's key is (tile << 32) | float_bits(view_z) with the tile
                                 given by the range the entry lies in */
  size_t bin_total;
  /* image workspace: tiles and pixels of the views back to back */
  size_t img_hdr;             /* uint32[4]: word 1 = num_rendered */
  size_t img_sub_bins;        /* S: sub-counters per tile (a number, not an offset) */
  size_t img_tile_counts;     /* uint32[T_total][S] instances per tile, split by Gaussian id mod S */
  size_t img_ranges;          /* uint2[T_total] */
  size_t img_state;           /* float4[N_total]: final_T, w_z_total, w_z2_total, bits(n_contrib) */
  size_t img_total;
} GftWorkspaceLayout;

void gft_workspace_layout_views(int P, int R, int n_views, const int* widths, const int* heights,
                                GftWorkspaceLayout* out);
void gft_workspace_layout(int P, int R, int width, int height, GftWorkspaceLayout* out);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (no counterpart in the reference, which has no profiler hooks: SURVEY §5).
 * gft_profile_enable(1) makes the calling thread's next gft_forward / gft_backward bracket every
 * stage with CUDA events on the caller's stream; gft_profile_read() then waits for them and
 * returns per-stage milliseconds (and static stage names) in launch order, at most `cap`.
 * gft_launch_count() is the process-wide number of kernels this library has launched.
 * ---------------------------------------------------------------------------------------- */
void gft_profile_enable(int on);
int gft_profile_read(float* ms, const char** names, int cap);
unsigned long long gft_launch_count(void);

/* Tunables / A-B switches of the kernels ("sort_cap", "sort_radix", "sort_match", "sub_bins",
 * "sort_adapt", "tile_order", "bwd_pred", "bwd_ring", "pbwd_minb", "pfwd_minb", "blend_half", "no_cull"; defaults
 * come from the environment variables GFT_SORT_CAP, ... read once).  Returns the previous value,
 * <0 for an unknown name.  Results never depend on them, only speed. */
int gft_set_option(const char* name, int value);

const char* gft_last_error(void);
int gft_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GFTORF_H_INCLUDED */
