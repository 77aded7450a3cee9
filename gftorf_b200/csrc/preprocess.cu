// preprocess.cu — per-Gaussian forward preprocessing for sm_100a, fused with the tile-count scan.
//
// Replaces, in ONE kernel:
//   - preprocessCUDA<3,7,2>   (cuda_rasterizer/forward.cu:251-419)  frustum-z cull, projection,
//     cov3D, EWA cov2D (+0.3), conic, radius, tile rectangle, SH->RGB, SH->(phase, amp),
//     phasor/quad synthesis with 1/d^2 falloff, distance-to-light, ndc distance;
//   - cub::DeviceScan::InclusiveSum over tiles_touched (rasterizer_impl.cu:307) — done here as a
//     single-pass chained scan with decoupled look-back, so point_offsets and num_rendered come
//     out of the same launch;
//   - the zero-fill of `pixels` and of the tile ranges (rasterize_points.cu:83,
//     rasterizer_impl.cu:341).
//
// Data layout written (HBM): one 80-byte "blend record" per Gaussian
//     x y ex ey | conA conB conC opacity | r g b dist | ph0 ph1 ph2 ph3 | ph4 ph5 ph6 ndc
// so the blend kernels fetch a Gaussian with five 16-byte loads from one place instead of seven
// scattered arrays; (ex, ey) are conservative half-extents of the region where the Gaussian can
// reach alpha >= 1/255, used for exact sub-tile culling.
#include <cstdio>
#include "common.cuh"
#include "kernels.h"

namespace gft {

namespace {

__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
  // Basis values multiplying sh[1..15] (forward.cu:37-59), in the dataflow of the reference
  // binary (SASS of preprocessCUDA, sm_100a): every product rounded, except the four polynomial
  // terms the compiler fused: 3xx-yy, 2zz-3xx-3yy (two FMAs), xx-3yy.  The phase SH term is
  // recovered by cancelling a large DC value, so one ulp here is visible in the phasor image.
  if (deg > 0) {
    b[1] = -__fmul_rn(kSH_C1, y);
    b[2] = __fmul_rn(kSH_C1, z);
    b[3] = -__fmul_rn(kSH_C1, x);
    if (deg > 1) {
      const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
      const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
      const float zz2 = __fadd_rn(zz, zz);
      const float xx_yy = __fsub_rn(xx, yy);
      b[4] = __fmul_rn(xy, kSH_C2_0);
      b[5] = __fmul_rn(yz, kSH_C2_1);
      b[6] = __fmul_rn(__fsub_rn(__fsub_rn(zz2, xx), yy), kSH_C2_2);
      b[7] = __fmul_rn(xz, kSH_C2_3);
      b[8] = __fmul_rn(xx_yy, kSH_C2_4);
      if (deg > 2) {
        const float t11 = __fsub_rn(__fmaf_rn(zz, 4.f, -xx), yy);
        b[9] = __fmul_rn(__fmul_rn(y, kSH_C3_0), __fmaf_rn(xx, 3.f, -yy));
        b[10] = __fmul_rn(__fmul_rn(xy, kSH_C3_1), z);
        b[11] = __fmul_rn(__fmul_rn(y, kSH_C3_2), t11);
        b[12] = __fmul_rn(__fmul_rn(z, kSH_C3_3), __fmaf_rn(yy, -3.f, __fmaf_rn(xx, -3.f, zz2)));
        b[13] = __fmul_rn(t11, __fmul_rn(x, kSH_C3_4));
        b[14] = __fmul_rn(xx_yy, __fmul_rn(z, kSH_C3_5));
        b[15] = __fmul_rn(__fmul_rn(x, kSH_C3_6), __fmaf_rn(yy, -3.f, xx));
      }
    }
  }
}

}  // namespace

namespace {

// SH -> RGB before the +0.5 (forward.cu:30-62); `sh` may point to global or shared memory
__device__ __forceinline__ void eval_sh_rgb(int D, const float* sh, const float* basis, float& r0,
                                            float& r1, float& r2) {
  r0 = __fmul_rn(kSH_C0, sh[0]); r1 = __fmul_rn(kSH_C0, sh[1]); r2 = __fmul_rn(kSH_C0, sh[2]);
  if (D > 0) {
    // result - C1*y*sh1 + C1*z*sh2 - C1*x*sh3   (basis carries the sign)
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      r0 = __fmaf_rn(basis[k], sh[3 * k + 0], r0);
      r1 = __fmaf_rn(basis[k], sh[3 * k + 1], r1);
      r2 = __fmaf_rn(basis[k], sh[3 * k + 2], r2);
    }
    if (D > 1) {
#pragma unroll
      for (int k = 4; k < 9; ++k) {
        r0 = __fmaf_rn(basis[k], sh[3 * k + 0], r0);
        r1 = __fmaf_rn(basis[k], sh[3 * k + 1], r1);
        r2 = __fmaf_rn(basis[k], sh[3 * k + 2], r2);
      }
      if (D > 2) {
#pragma unroll
        for (int k = 9; k < 16; ++k) {
          r0 = __fmaf_rn(basis[k], sh[3 * k + 0], r0);
          r1 = __fmaf_rn(basis[k], sh[3 * k + 1], r1);
          r2 = __fmaf_rn(basis[k], sh[3 * k + 2], r2);
        }
      }
    }
  }
}

// SH -> (phase, amplitude) before the +0.5 (forward.cu:79-111)
__device__ __forceinline__ void eval_sh_pa(int D, const float* sp, const float* basis, float& q0,
                                           float& q1) {
  q0 = __fmul_rn(kSH_C0, sp[0]); q1 = __fmul_rn(kSH_C0, sp[1]);
  if (D > 0) {
#pragma unroll
    for (int k = 1; k < 4; ++k) { q0 = __fmaf_rn(basis[k], sp[2 * k], q0); q1 = __fmaf_rn(basis[k], sp[2 * k + 1], q1); }
    if (D > 1) {
#pragma unroll
      for (int k = 4; k < 9; ++k) { q0 = __fmaf_rn(basis[k], sp[2 * k], q0); q1 = __fmaf_rn(basis[k], sp[2 * k + 1], q1); }
      if (D > 2) {
#pragma unroll
        for (int k = 9; k < 16; ++k) { q0 = __fmaf_rn(basis[k], sp[2 * k], q0); q1 = __fmaf_rn(basis[k], sp[2 * k + 1], q1); }
      }
    }
  }
}

}  // namespace

// One block = 256 consecutive Gaussians, ALL views of the batch.  The per-Gaussian parameters
// (mean, scale, rotation, opacity and the 48 + 32 SH floats: 364 B) are read once; per view the
// kernel culls, projects, evaluates colour and phasor, writes the view's blend record and counts
// the Gaussian into every tile it touches (tile_counts, the input of the tile-segmented binning).
//   pass A (colour SH rows of the warp staged in shared memory): geometry + colour per view
//   pass B (phase/amplitude SH rows staged in the same buffer):   phasor per view
// The arithmetic of every view is the single-view sequence pinned to the reference binary.
template <int MINB>
__global__ void __launch_bounds__(GFT_BLOCK, MINB)
preprocess_fwd_kernel(const __grid_constant__ PreprocessParams p) {
  extern __shared__ float fwd_stage[];  // GFT_STAGE_FLOATS_PER_WARP floats per warp

  const int idx = (int)(blockIdx.x * GFT_BLOCK + threadIdx.x);
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const bool in_range = idx < p.P;
  const size_t P = (size_t)p.P;

  float px = 0.f, py = 0.f, pz = 0.f;
  if (in_range) {
    px = __ldg(p.means3D + 3 * idx + 0);
    py = __ldg(p.means3D + 3 * idx + 1);
    pz = __ldg(p.means3D + 3 * idx + 2);
  }

  // ---- frustum test of every view (in_frustum, auxiliary.h:152-179: z only, NaN passes) ----------
  uint32_t alive_mask = 0;
  if (in_range) {
    for (int v = 0; v < p.nviews; ++v) {
      const ViewCam& vc = p.views[v];
      const float vz = xform_row(vc.viewmatrix, 2, px, py, pz);
      const bool ok = !(vz < vc.near_n || vz > vc.far_n);
      if (!ok && p.prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
      alive_mask |= ok ? (1u << v) : 0u;
    }
  }

  // ---- view-independent: 3D covariance (forward.cu:172-206), saved for the backward -------------
  Cov3 c3 = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float opacity = 0.f;
  if (alive_mask) {
    if (p.cov3D_precomp != nullptr) {
      const float* c = p.cov3D_precomp + 6 * (size_t)idx;
      c3.c0 = __ldg(c + 0); c3.c1 = __ldg(c + 1); c3.c2 = __ldg(c + 2);
      c3.c3 = __ldg(c + 3); c3.c4 = __ldg(c + 4); c3.c5 = __ldg(c + 5);
    } else {
      const float* s = p.scales + 3 * (size_t)idx;
      const float4 q = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
      c3 = cov3d_from_scale_rot(__ldg(s), __ldg(s + 1), __ldg(s + 2), p.scale_modifier, q.x, q.y,
                                q.z, q.w);
      float2* dst = reinterpret_cast<float2*>(p.g.cov3D + 6 * (size_t)idx);
      dst[0] = make_float2(c3.c0, c3.c1);
      dst[1] = make_float2(c3.c2, c3.c3);
      dst[2] = make_float2(c3.c4, c3.c5);
    }
    opacity = __ldg(p.opacities + idx);
  }

  float* wbuf = fwd_stage + warp * GFT_STAGE_FLOATS_PER_WARP;
  const int wfirst = (int)(blockIdx.x * GFT_BLOCK + warp * 32);
  const int nrows = max(0, min(32, p.P - wfirst));
  const bool any_alive = __any_sync(0xffffffffu, alive_mask != 0u);
  const bool st_sh = p.shs != nullptr && p.M == 16 && any_alive;
  const bool st_shp = p.shs_p != nullptr && p.M_p == 16 && any_alive;
  if (st_sh) {
    warp_stage_in<48>(p.shs + (size_t)wfirst * 48, nrows, wbuf, lane);
    __syncwarp();
  }

  // ================= pass A: geometry + colour, per view ========================================
  for (int v = 0; v < p.nviews; ++v) {
    const ViewCam& vc = p.views[v];
    const float* __restrict__ V = vc.viewmatrix;
    const float* __restrict__ PM = vc.projmatrix;
    bool alive = (alive_mask >> v) & 1u;
    uint32_t tiles = 0;
    int radius_i = 0;
    uint32_t rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;
    float vx = 0.f, vy = 0.f, vz = 0.f, pix_x = 0.f, pix_y = 0.f;
    float3 cov = make_float3(0.f, 0.f, 0.f);
    float det = 0.f;

    if (alive) {
      vz = xform_row(V, 2, px, py, pz);
      vx = xform_row(V, 0, px, py, pz);
      vy = xform_row(V, 1, px, py, pz);
      const float hx = xform_row(PM, 0, px, py, pz);
      const float hy = xform_row(PM, 1, px, py, pz);
      const float hw = xform_row(PM, 3, px, py, pz);
      const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
      const float projx = __fmul_rn(hx, p_w);
      const float projy = __fmul_rn(hy, p_w);

      const Tmat T = ewa_T(V, vx, vy, vz, vc.focal_x, vc.focal_y, vc.tan_fovx, vc.tan_fovy);
      cov = ewa_cov2d(T, c3);
      det = __fmaf_rn(cov.x, cov.z, -__fmul_rn(cov.y, cov.y));
      alive = !(det == 0.0f);  // forward.cu:325
      if (alive) {
        const float mid = __fmul_rn(__fadd_rn(cov.x, cov.z), 0.5f);
        const float disc = fmaxf(__fmaf_rn(mid, mid, -det), 0.1f);
        const float s = __fsqrt_rn(disc);
        const float lam = fmaxf(__fadd_rn(mid, s), __fsub_rn(mid, s));
        const float my_radius = ceilf(__fmul_rn(__fsqrt_rn(lam), 3.f));
        radius_i = (int)my_radius;
        pix_x = ndc2pix(projx, vc.W);
        pix_y = ndc2pix(projy, vc.H);
        tile_rect(pix_x, pix_y, radius_i, vc.grid_x, vc.grid_y, rx0, ry0, rx1, ry1);
        tiles = (rx1 - rx0) * (ry1 - ry0);
        alive = tiles != 0;
      }
    }
    if (!alive) { tiles = 0; radius_i = 0; alive_mask &= ~(1u << v); }
    if (in_range) {
      vc.radii[idx] = radius_i;
      vc.pixels[idx] = 0.f;
      p.g.tiles_touched[v * P + idx] = tiles;
    }

    // one count per (Gaussian, tile) instance of this view
    {
      const uint32_t S = (uint32_t)p.sub_bins;
      uint32_t* cnt = p.tile_counts + (size_t)vc.tile_base * S;
      const uint32_t gx = (uint32_t)vc.grid_x;
      for_each_tile(rx0, ry0, rx1, tiles, (uint32_t)idx & (S - 1u), lane,
                    [&](uint32_t tx, uint32_t ty, uint32_t sub) { atomicAdd(cnt + (ty * gx + tx) * S + sub, 1u); });
    }

    if (alive) {
      // ---- colour (forward.cu:346-359) -------------------------------------------------------
      float cr = 0.f, cg = 0.f, cb = 0.f;
      uint32_t clamp_bits = 0;
      if (p.colors_precomp != nullptr) {
        cr = __ldg(p.colors_precomp + 3 * (size_t)idx + 0);
        cg = __ldg(p.colors_precomp + 3 * (size_t)idx + 1);
        cb = __ldg(p.colors_precomp + 3 * (size_t)idx + 2);
      }
      if (p.shs != nullptr) {   // SH overwrites a precomputed colour when both are given (forward.cu:346-359)
        // view direction and SH basis (forward.cu:25-27): FMUL(dy,dy) -> FFMA(dx,dx,.) ->
        // FFMA(dz,dz,.), IEEE sqrt/div, as in the reference binary
        float basis[16];
        float dirx = __fsub_rn(px, __ldg(vc.campos + 0));
        float diry = __fsub_rn(py, __ldg(vc.campos + 1));
        float dirz = __fsub_rn(pz, __ldg(vc.campos + 2));
        const float len = __fsqrt_rn(dot3c(dirx, dirx, diry, diry, dirz, dirz));
        dirx = __fdiv_rn(dirx, len);
        diry = __fdiv_rn(diry, len);
        dirz = __fdiv_rn(dirz, len);
        sh_basis(p.D, dirx, diry, dirz, basis);
        float r0, r1, r2;
        if (st_sh) eval_sh_rgb(p.D, wbuf + lane * 49, basis, r0, r1, r2);
        else eval_sh_rgb(p.D, p.shs + (size_t)idx * p.M * 3, basis, r0, r1, r2);
        r0 = __fadd_rn(r0, 0.5f); r1 = __fadd_rn(r1, 0.5f); r2 = __fadd_rn(r2, 0.5f);
        clamp_bits |= (r0 < 0.f) ? 1u : 0u;
        clamp_bits |= (r1 < 0.f) ? (1u << 8) : 0u;
        clamp_bits |= (r2 < 0.f) ? (1u << 16) : 0u;
        cr = fmaxf(r0, 0.f); cg = fmaxf(r1, 0.f); cb = fmaxf(r2, 0.f);
      }

      const float det_inv = __frcp_rn(det);
      const float conA = __fmul_rn(cov.z, det_inv);
      const float conB = __fmul_rn(cov.y, -det_inv);
      const float conC = __fmul_rn(cov.x, det_inv);
      // distance to the light (forward.cu:361): y^2 rounded, x^2 and z^2 fused, as compiled
      const float dist = __fsqrt_rn(dot3c(vx, vx, vy, vy, vz, vz));

      // ---- conservative contribution extents for sub-tile culling ---------------------------
      // A pixel can pass the alpha test only if opacity*exp(power) >= 1/255, i.e.
      // d^T Q d <= 2 ln(255*opacity).  The rounding error of the float `power` chain is bounded
      // by 2k*cond(Sigma)*q (k = 4e-7), so inflating the threshold by 1/(1-4k*cond) keeps the box
      // a superset of what the reference's arithmetic can accept.  Anything doubtful -> no cull.
      float ex = __int_as_float(0x7f800000), ey = __int_as_float(0x7f800000);  // +inf
      if (!p.subtile_cull) {
        // keep +inf: no culling
      } else if (opacity < (1.0f / 255.0f)) {
        ex = ey = -__int_as_float(0x7f800000);  // alpha <= opacity < 1/255: can never contribute
      } else {
        const double dA = (double)conA, dB = (double)conB, dC = (double)conC;
        const double dq = dA * dC - dB * dB;
        const float mid = 0.5f * (cov.x + cov.z);
        const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
        const float lmax = mid + sq, lmin = mid - sq;
        const float cond = lmax / lmin;
        if (dq > 0.0 && dA > 0.0 && dC > 0.0 && lmin > 0.f && cond < 2.0e5f && cond == cond) {
          const float t2 = (2.0f * __logf(255.0f * opacity) + 0.04f) / (1.0f - 1.6e-6f * cond);
          const float sxx = (float)(dC / dq), syy = (float)(dA / dq);
          ex = sqrtf(t2 * sxx) * 1.0001f + 0.02f;
          ey = sqrtf(t2 * syy) * 1.0001f + 0.02f;
        }
      }

      float4* rec = reinterpret_cast<float4*>(p.g.rec + (v * P + idx) * GFT_REC_FLOATS);
      rec[0] = make_float4(pix_x, pix_y, ex, ey);
      rec[1] = make_float4(conA, conB, conC, opacity);
      rec[2] = make_float4(cr, cg, cb, dist);
      p.g.depths[v * P + idx] = vz;
      p.g.clamped[v * P + idx] = clamp_bits;
      reinterpret_cast<uint2*>(p.g.rect)[v * P + idx] = make_uint2(rx0 | (ry0 << 16), rx1 | (ry1 << 16));
    }
  }

  __syncwarp();
  if (st_shp) {
    warp_stage_in<32>(p.shs_p + (size_t)wfirst * 32, nrows, wbuf, lane);
    __syncwarp();
  }

  // ================= pass B: phasor, per view (forward.cu:361-407) ==============================
  for (int v = 0; v < p.nviews; ++v) {
    if (!((alive_mask >> v) & 1u)) continue;
    const ViewCam& vc = p.views[v];
    const float* __restrict__ V = vc.viewmatrix;
    // the same three rows as in pass A: identical bits
    const float vz = xform_row(V, 2, px, py, pz);
    const float vx = xform_row(V, 0, px, py, pz);
    const float vy = xform_row(V, 1, px, py, pz);
    const float dist = __fsqrt_rn(dot3c(vx, vx, vy, vy, vz, vz));
    const float ndc = __fmul_rn(__fsub_rn(1.f, __fdiv_rn(vc.near_n, dist)),
                                __fdiv_rn(vc.far_n, __fsub_rn(vc.far_n, vc.near_n)));
    const float factor = __frcp_rn(__fmul_rn(dist, dist));

    // With neither shs_p nor phasors_precomp the reference leaves real_img_amp uninitialised
    // (SURVEY A.7-7); we define those features as 0.
    float ph[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float pa0 = 0.f, pa1 = 0.f;
    bool have_ph = false;
    float phase = 0.f, amp = 0.f;
    uint32_t amp_clamped = 0;
    if (p.phasors_precomp != nullptr) {
      phase = __fmul_rn(dist, vc.dist2phase);
      pa0 = __ldg(p.phasors_precomp + 2 * (size_t)idx + 0);
      pa1 = __ldg(p.phasors_precomp + 2 * (size_t)idx + 1);
      if (vc.use_view_dependent_phase) phase = __fadd_rn(phase, pa0);
      amp = pa1;
      have_ph = true;
    }
    if (p.shs_p != nullptr) {
      float basis[16];
      float dirx = __fsub_rn(px, __ldg(vc.campos + 0));
      float diry = __fsub_rn(py, __ldg(vc.campos + 1));
      float dirz = __fsub_rn(pz, __ldg(vc.campos + 2));
      const float len = __fsqrt_rn(dot3c(dirx, dirx, diry, diry, dirz, dirz));
      dirx = __fdiv_rn(dirx, len);
      diry = __fdiv_rn(diry, len);
      dirz = __fdiv_rn(dirz, len);
      sh_basis(p.D, dirx, diry, dirz, basis);
      float q0, q1, dc0;
      if (st_shp) {
        eval_sh_pa(p.D, wbuf + lane * 33, basis, q0, q1);
        dc0 = wbuf[lane * 33];
      } else {
        const float* sp = p.shs_p + (size_t)idx * p.M_p * 2;
        eval_sh_pa(p.D, sp, basis, q0, q1);
        dc0 = sp[0];
      }
      q1 = __fadd_rn(q1, 0.5f);
      // remove phase DC (forward.cu:115): ((q0 + 0.5) - 0.5) - FMUL(C0, sh_p[0].x)
      q0 = __fsub_rn(__fsub_rn(__fadd_rn(q0, 0.5f), 0.5f), __fmul_rn(kSH_C0, dc0));
      if (q1 < 0.f) { amp_clamped = 1u << 24; q1 = 0.f; }
      pa0 = q0; pa1 = q1;
      phase = __fmaf_rn(dist, vc.dist2phase, vc.phase_offset);
      if (vc.use_view_dependent_phase) phase = __fadd_rn(q0, phase);
      amp = q1;
      have_ph = true;
    }
    if (have_ph) {
      float sn, cs;
      sincosf(phase, &sn, &cs);
      const float dc = vc.dc_offset;   // (trig +- dc) * amp * factor, left to right (forward.cu:399-406)
      ph[0] = __fmul_rn(__fmul_rn(cs, amp), factor);
      ph[1] = __fmul_rn(__fmul_rn(sn, amp), factor);
      ph[2] = __fmul_rn(amp, factor);
      ph[3] = __fmul_rn(__fmul_rn(__fadd_rn(cs, dc), amp), factor);
      ph[4] = __fmul_rn(__fmul_rn(__fadd_rn(-cs, dc), amp), factor);
      ph[5] = __fmul_rn(__fmul_rn(__fadd_rn(sn, dc), amp), factor);
      ph[6] = __fmul_rn(__fmul_rn(__fadd_rn(-sn, dc), amp), factor);
    }
    float4* rec = reinterpret_cast<float4*>(p.g.rec + (v * P + idx) * GFT_REC_FLOATS);
    rec[3] = make_float4(ph[0], ph[1], ph[2], ph[3]);
    rec[4] = make_float4(ph[4], ph[5], ph[6], ndc);
    if (amp_clamped) p.g.clamped[v * P + idx] |= amp_clamped;   // own word, written in pass A
    reinterpret_cast<float2*>(p.g.pa)[v * P + idx] = make_float2(pa0, pa1);
  }
}

// markVisible / checkFrustum, rasterizer_impl.cu:54-68
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D,
                                    const float* __restrict__ V, uint8_t* __restrict__ present,
                                    float near_n, float far_n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float vz = xform_row(V, 2, __ldg(means3D + 3 * idx), __ldg(means3D + 3 * idx + 1),
                             __ldg(means3D + 3 * idx + 2));
  present[idx] = !(vz < near_n || vz > far_n);
}

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t stream) {
  const int blocks = (p.P + GFT_BLOCK - 1) / GFT_BLOCK;
  const int smem = (GFT_BLOCK / 32) * GFT_STAGE_FLOATS_PER_WARP * (int)sizeof(float);
  // 3 resident blocks per SM (79 registers, no spills; the unbounded build took 83 and ran 2 per SM:
  // ncu r2_g) or 4 (64 registers, 72 B spilled): option pfwd_minb.  The carveout preference lets
  // 4 x 49 KB of staging fit.
  static unsigned long long ok3 = 0, ok4 = 0;
  auto go = [&](auto kernel, unsigned long long* ok) {
    if (!(*ok & (1ull << 63))) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      __atomic_fetch_or(ok, 1ull << 63, __ATOMIC_RELAXED);
    }
    ensure_dynamic_smem(kernel, smem, ok);
    kernel<<<blocks, GFT_BLOCK, smem, stream>>>(p);
  };
  if (option(OPT_PFWD_MINB) == 4) go(preprocess_fwd_kernel<4>, &ok4);
  else go(preprocess_fwd_kernel<3>, &ok3);
  note_launches(1);
}

void launch_mark_visible(int P, const float* means3D, const float* V, uint8_t* present,
                         float near_n, float far_n, cudaStream_t stream) {
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, V, present, near_n, far_n);
  note_launches(1);
}

}  // namespace gft
