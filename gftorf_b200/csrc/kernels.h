// kernels.h — internal launch interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

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

// Records the thread-local error string behind gft_last_error() and returns `code` (api.cu).
int set_error(int code, const char* msg);

// Warps per blend block (8 = a whole 16x16 tile, 4 = a 16x8 half, 2 = a 16x4 quarter); see
// blend_fwd.cu.  GFT_BLEND_WARPS overrides the choice (for A/B measurements).
int blend_block_warps(int tiles);

struct PreprocessParams {
  int P, D, M, M_p;
  int W, H, grid_x, grid_y, num_tiles;
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
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int prefiltered;
  float near_n, far_n, dist2phase;
  int use_view_dependent_phase;
  float phase_offset, dc_offset;
  int subtile_cull;       // 0: write infinite extents (blend kernels then test every pair)
  // outputs
  int* radii;
  float* pixels;
  float* rec;             // [P][20]
  float* depths;          // [P]
  uint32_t* tiles_touched;
  uint32_t* point_offsets;
  uint16_t* rect;         // [P][4]
  float* cov3D;           // [P][6]
  uint32_t* clamped;      // [P] packed r|g<<8|b<<16|amp<<24
  float* pa;              // [P][2]
  uint2* ranges;          // [T]
  // scan
  uint32_t* scan_ticket;   // header word 0 (unused since block ids are static)
  uint32_t* key_format_out; // header words 2,3: depth_bits, depth_base (for the debug accessor)
  int key_depth_bits; uint32_t key_depth_base;
  unsigned long long* scan_state;
  uint32_t* num_rendered;
};

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* V, uint8_t* present,
                         float near_n, float far_n, cudaStream_t stream);

// ---- binning ---------------------------------------------------------------------------------
// Sort-key format.  The reference's key is (tile << 32) | float_bits(view_z).  Visible Gaussians
// have near <= view_z <= far (the frustum test), and positive floats order like their bit
// patterns, so (float_bits(view_z) - float_bits(near)) needs only depth_bits =
// bit_length(float_bits(far) - float_bits(near)) bits and sorts identically: the compact key is
// (tile << depth_bits) | (float_bits(view_z) - depth_base), typically 37-39 significant bits
// instead of 43-45 — one radix pass fewer.  depth_bits = 32, depth_base = 0 is the reference's format.
struct KeyFormat {
  int depth_bits;
  uint32_t depth_base;
};
KeyFormat key_format(float near_n, float far_n);

void launch_duplicate_keys(int P, const int* radii, const uint16_t* rect, const float* depths,
                           const uint32_t* point_offsets, uint64_t* keys, uint32_t* values,
                           int grid_x, uint32_t capacity, KeyFormat kf, cudaStream_t stream);
// Stable LSD radix sort of (key,value) pairs on bits [0, end_bit).  Returns temp bytes needed
// when d_temp == nullptr.
size_t sort_pairs_temp_bytes(int R);
// R: capacity (grid / temp sizing); d_R: optional device pointer to the actual count (<= R)
int sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
               const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
               cudaStream_t stream, const uint32_t* d_R = nullptr);
bool sort_result_in_out(int end_bit);
void launch_identify_ranges(int R, const uint32_t* d_R, const uint64_t* keys, uint2* ranges,
                            KeyFormat kf, cudaStream_t stream);

// ---- blend -----------------------------------------------------------------------------------
struct BlendFwdParams {
  int W, H, grid_x, grid_y;
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* rec;
  const float* bg;
  int bg_mode;
  float4* img_state;  // final_T, w_z_total, w_z2_total, n_contrib bits
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
void launch_blend_fwd(const BlendFwdParams& p, cudaStream_t stream);

struct BlendBwdParams {
  int W, H, grid_x, grid_y;
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* rec;
  const float* bg;
  int bg_mode;
  const float4* img_state;
  const float* dL_dcolor;
  const float* dL_dphasor;
  const float* dL_ddepth;
  const float* dL_dacc;
  const float* dL_ddd;
  float* grad_rec;  // [P][16], zero-filled before launch
};
void launch_blend_bwd(const BlendBwdParams& p, cudaStream_t stream);

// ---- preprocess backward ---------------------------------------------------------------------
struct PreprocessBwdParams {
  int P, D, M, M_p;
  int W, H;
  const float* means3D;
  const int* radii;
  const float* shs;
  const float* shs_p;
  const uint32_t* clamped;
  const float* scales;
  const float* rotations;
  float scale_modifier;
  const float* cov3D;  // precomp or saved
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float focal_x, focal_y, tan_fovx, tan_fovy;
  const float* rec;       // forward blend records (dist)
  const float* pa;        // [P][2]
  const float* grad_rec;  // [P][16] from blend backward
  float near_n, far_n, dist2phase;
  int use_view_dependent_phase;
  float phase_offset, dc_offset;
  int accumulate;  // add into the parameter gradients instead of overwriting
  // outputs
  float* dL_dmeans2D;
  float* dL_dopacity;
  float* dL_dmeans3D;
  float* dL_dsh;
  float* dL_dsh_p;
  float* dL_dscales;
  float* dL_drotations;
  float* dL_dphase_offset;
  float* dL_ddc_offset;
  float* dL_dcolors;
  float* dL_dcov3D;
  float* dL_dconic;
  float* dL_ddist;
  float* dL_dndc;
};
void launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream);

// ---- knn -------------------------------------------------------------------------------------
size_t knn_workspace_bytes(int P);
int knn_dist2(const float* points, int P, float* out, char* workspace, cudaStream_t stream);

}  // namespace gft
