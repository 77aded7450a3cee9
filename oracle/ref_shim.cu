// ref_shim.cu — TEST INFRASTRUCTURE ONLY.  A C-ABI door onto the UNMODIFIED reference kernels.
//
// This file contains no rasterizer arithmetic.  It is linked with the reference's own translation
// units, compiled where they lie under /root/reference (see oracle/Makefile), into
// oracle/_ref/libgftorf_ref.so:
//   submodules/diff-gaussian-rasterization-w-tof/cuda_rasterizer/{rasterizer_impl,forward,backward}.cu
//   submodules/simple-knn/simple_knn.cu
// and forwards the same POD argument structs our product library takes (include/gftorf.h) to
//   CudaRasterizer::Rasterizer::forward / backward / markVisible   (cuda_rasterizer/rasterizer.h:20-105)
//   SimpleKNN::knn                                                 (simple-knn/simple_knn.h:15-19)
// It stands where the reference's torch binding stands (rasterize_points.cu:35-304, spatial.cu:15-26),
// minus torch: the caller (tests / bench reference arm) allocates and zero-fills exactly as the
// binding does (rasterize_points.cu:80-101,222-236).
//
// Only tests/, __graft_entry__.smoke() and bench.py's reference/baseline legs may load the result.
#include <cuda_runtime.h>
#include <functional>
#include <cstdint>
#include <cstring>
#include <string>
#include <stdexcept>

#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"
#include "simple_knn.h"
#include "gftorf.h"

static thread_local std::string g_err;

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

// The reference launches on the legacy default stream; `stream` is accepted for signature parity
// and must be 0 / the legacy stream (the caller synchronises around the call otherwise).
int ref_forward(const GftForwardArgs* a, gft_alloc_fn geom_alloc, gft_alloc_fn binning_alloc,
                gft_alloc_fn img_alloc, void* ctx, gft_stream_t /*stream*/) {
  try {
    if (a->P == 0) return 0;  // rasterize_points.cu:104
    std::function<char*(size_t)> geomF = [&](size_t n) { return geom_alloc(ctx, n); };
    std::function<char*(size_t)> binF = [&](size_t n) { return binning_alloc(ctx, n); };
    std::function<char*(size_t)> imgF = [&](size_t n) { return img_alloc(ctx, n); };
    int R = CudaRasterizer::Rasterizer::forward(
        geomF, binF, imgF, a->P, a->sh_degree, a->M, a->M_p, a->background, a->width, a->height,
        a->means3D, a->shs, a->shs_p, a->colors_precomp, a->phasors_precomp, a->opacities,
        a->scales, a->scale_modifier, a->rotations, a->cov3D_precomp, a->viewmatrix,
        a->projmatrix, a->campos, a->tan_fovx, a->tan_fovy, a->prefiltered != 0, a->out_color,
        a->out_phasor, a->out_depth, a->out_normal, a->out_acc, a->out_entropy,
        a->out_depth_distortion, a->out_amp_distortion, a->pixels, a->out_distribution, a->radii,
        a->debug != 0, a->near_n, a->far_n, a->depth_range, a->use_view_dependent_phase != 0,
        a->phase_offset, a->dc_offset);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return -1; }
    return R;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}

// Extra pointers the reference backward wants that are not in GftBackwardArgs (its always-
// materialised intermediates) come through GftBackwardArgs' optional fields: dL_dcolors,
// dL_dphasors, dL_dcov3D, dL_dconic, dL_ddist, dL_dndc must all be non-NULL and zero-filled here,
// as must every required output (the reference accumulates with atomics into zeroed tensors).
int ref_backward(const GftBackwardArgs* a, gft_stream_t /*stream*/) {
  try {
    if (a->P == 0) return 0;  // rasterize_points.cu:238
    CudaRasterizer::Rasterizer::backward(
        a->P, a->sh_degree, a->M, a->M_p, a->R, a->background, a->width, a->height, a->means3D,
        a->shs, a->shs_p, a->colors_precomp, a->phasors_precomp, a->scales, a->scale_modifier,
        a->rotations, a->cov3D_precomp, a->viewmatrix, a->projmatrix, a->campos, a->tan_fovx,
        a->tan_fovy, a->radii, const_cast<char*>(a->geom_buffer),
        const_cast<char*>(a->binning_buffer), const_cast<char*>(a->img_buffer), a->dL_dout_color,
        a->dL_dout_phasor, a->dL_dout_depth, /*dL_dpix_n*/ nullptr, a->dL_dout_acc,
        /*dL_dpix_e*/ nullptr, a->dL_dout_depth_distortion, /*dL_dpix_ad*/ nullptr,
        a->dL_dmeans2D, a->dL_dconic, a->dL_dopacity, a->dL_dcolors, a->dL_dphasors, a->dL_ddist,
        a->dL_dndc, a->dL_dmeans3D, a->dL_dcov3D, a->dL_dsh, a->dL_dsh_p, a->dL_dscales,
        a->dL_drotations, a->dL_dphase_offset, a->dL_ddc_offset, a->debug != 0, a->near_n,
        a->far_n, a->depth_range, a->use_view_dependent_phase != 0, a->phase_offset,
        a->dc_offset);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return -1; }
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}

int ref_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, float near_n, float far_n, gft_stream_t /*stream*/) {
  if (P == 0) return 0;
  CudaRasterizer::Rasterizer::markVisible(P, const_cast<float*>(means3D),
                                          const_cast<float*>(viewmatrix),
                                          const_cast<float*>(projmatrix),
                                          reinterpret_cast<bool*>(present), near_n, far_n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return -1; }
  return 0;
}

int ref_dist2(const float* points, int P, float* out, char* /*workspace*/, gft_stream_t /*stream*/) {
  try {
    SimpleKNN::knn(P, reinterpret_cast<float3*>(const_cast<float*>(points)), out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return -1; }
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}

// Byte offsets of the arrays inside the reference's three opaque buffers, obtained by running the
// reference's OWN fromChunk walk (rasterizer_impl.cu:161-211) over the caller's base pointers —
// nothing about the layout is re-derived here.  (SURVEY Appendix B.)
typedef struct RefLayout {
  size_t geom_depths, geom_ndc, geom_clamped, geom_clamped_p, geom_radii, geom_means2D, geom_cov3D,
      geom_conic_opacity, geom_rgb, geom_real_img_amp, geom_dists, geom_pa, geom_tiles_touched,
      geom_point_offsets, geom_total;
  size_t img_accum_alpha, img_w_z_total, img_w_z2_total, img_n_contrib, img_ranges, img_total;
  size_t bin_point_list, bin_point_list_unsorted, bin_keys, bin_keys_unsorted, bin_total;
} RefLayout;

void ref_layout(const char* geom_base, const char* img_base, const char* bin_base, int P, int R,
                int width, int height, RefLayout* o) {
  using namespace CudaRasterizer;
  std::memset(o, 0, sizeof(*o));
  auto off = [](const void* p, const char* base) { return (size_t)((const char*)p - base); };
  if (geom_base && P > 0) {
    char* c = const_cast<char*>(geom_base);
    GeometryState g = GeometryState::fromChunk(c, (size_t)P);
    o->geom_depths = off(g.depths, geom_base);
    o->geom_ndc = off(g.dists_to_light_ndc, geom_base);
    o->geom_clamped = off(g.clamped, geom_base);
    o->geom_clamped_p = off(g.clamped_p, geom_base);
    o->geom_radii = off(g.internal_radii, geom_base);
    o->geom_means2D = off(g.means2D, geom_base);
    o->geom_cov3D = off(g.cov3D, geom_base);
    o->geom_conic_opacity = off(g.conic_opacity, geom_base);
    o->geom_rgb = off(g.rgb, geom_base);
    o->geom_real_img_amp = off(g.real_img_amp, geom_base);
    o->geom_dists = off(g.dists_to_light, geom_base);
    o->geom_pa = off(g.phase_amplitude_from_sh, geom_base);
    o->geom_tiles_touched = off(g.tiles_touched, geom_base);
    o->geom_point_offsets = off(g.point_offsets, geom_base);
    o->geom_total = off(c, geom_base);
  }
  if (img_base && width > 0 && height > 0) {
    char* c = const_cast<char*>(img_base);
    ImageState im = ImageState::fromChunk(c, (size_t)width * height);
    o->img_accum_alpha = off(im.accum_alpha, img_base);
    o->img_w_z_total = off(im.w_z_total, img_base);
    o->img_w_z2_total = off(im.w_z2_total, img_base);
    o->img_n_contrib = off(im.n_contrib, img_base);
    o->img_ranges = off(im.ranges, img_base);
    o->img_total = off(c, img_base);
  }
  if (bin_base && R > 0) {
    char* c = const_cast<char*>(bin_base);
    BinningState b = BinningState::fromChunk(c, (size_t)R);
    o->bin_point_list = off(b.point_list, bin_base);
    o->bin_point_list_unsorted = off(b.point_list_unsorted, bin_base);
    o->bin_keys = off(b.point_list_keys, bin_base);
    o->bin_keys_unsorted = off(b.point_list_keys_unsorted, bin_base);
    o->bin_total = off(c, bin_base);
  }
}

}  // extern "C"
