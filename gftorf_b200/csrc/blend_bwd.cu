// blend_bwd.cu — back-to-front replay of the per-tile blend and its gradients, for sm_100a.
//
// Replaces renderCUDA<3,7> of the backward pass (cuda_rasterizer/backward.cu:609-889).
//
// The reference issues 18 scalar global atomicAdds for every contributing (pixel, Gaussian) pair
// (backward.cu:788,802,812,832,876-877,881-883,886).  Here:
//   - a 16x16 tile is walked by 8 warps of 8x4 pixels, 32 Gaussians at a time; a warp skips
//     Gaussians whose conservative alpha>=1/255 box misses its patch (exact, see blend_fwd.cu);
//   - the 18 per-pair partial gradients are summed over the warp's 32 pixels with a halving
//     butterfly (20 -> 10 -> 5 values per lane, then a 3-step butterfly on the 5: 30 shuffles
//     instead of 18*5 = 90), which leaves record floats [4g..4g+3] and [16+g] of the Gaussian's
//     20-float gradient record complete in the lanes of lane-group g = lane>>3;
//   - four lanes then issue one 16-byte vector reduction each (REDG.E.ADD.F32x4) plus two scalar
//     ones: 6 reductions per (warp, Gaussian) instead of up to 32*18.
//   - batches behind the furthest last-contributor of the tile are never loaded.
// The per-pixel recurrences are the reference's, with `accum_rec` advanced at the end of an
// iteration instead of the start of the next (same operations, same order, no last_color copy).
//
// Gradient record layout [P][20] (private; consumed by preprocess_bwd.cu):
//   0 mean2D.x  1 mean2D.y  2 conic.x  3 conic.y | 4 conic.w  5 opacity  6 col.r  7 col.g |
//   8 col.b  9 dist  10 ndc  11 ph0 | 12 ph1 13 ph2 14 ph3 15 ph4 | 16 ph5 17 ph6 18,19 unused
#include "common.cuh"
#include "kernels.h"

namespace gft {

namespace {

constexpr int BATCH = 256;

struct BwdSmem {
  float4 r0[BATCH];  // x y ex ey
  float4 r1[BATCH];  // conA conB conC opacity
  float4 r2[BATCH];  // r g b dist
  float4 r3[BATCH];  // ph0..ph3
  float4 r4[BATCH];  // ph4 ph5 ph6 ndc
  int id[BATCH];
  uint32_t wmax[GFT_BLOCK / 32];
};

// Butterfly value index -> which quantity.  Lane-group g ends up owning values 5g..5g+4 which are
// record floats 4g..4g+3 and 16+g.
//   v[0..3]   = rec 0..3      v[4]  = rec 16
//   v[5..8]   = rec 4..7      v[9]  = rec 17
//   v[10..13] = rec 8..11     v[14] = rec 18 (pad)
//   v[15..18] = rec 12..15    v[19] = rec 19 (pad)

}  // namespace

__global__ void __launch_bounds__(GFT_BLOCK, 2)
blend_bwd_kernel(BlendBwdParams p) {
  __shared__ BwdSmem s;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t tile_x = tile % (uint32_t)p.grid_x, tile_y = tile / (uint32_t)p.grid_x;
  const uint32_t px0 = tile_x * GFT_TILE_X + (warp & 1u) * 8u;
  const uint32_t py0 = tile_y * GFT_TILE_Y + (warp >> 1) * 4u;
  const uint32_t pix_x = px0 + (lane & 7u), pix_y = py0 + (lane >> 3);
  const bool inside = pix_x < (uint32_t)p.W && pix_y < (uint32_t)p.H;
  const uint32_t pix_id = (uint32_t)p.W * pix_y + pix_x;
  const float pixfx = (float)pix_x, pixfy = (float)pix_y;
  const float patch_x0 = (float)px0, patch_x1 = (float)(px0 + 7u);
  const float patch_y0 = (float)py0, patch_y1 = (float)(py0 + 3u);
  const size_t HW = (size_t)p.H * (size_t)p.W;

  const uint2 range = p.ranges[tile];
  const int n = (int)(range.y - range.x);

  // per-pixel forward state and incoming gradients
  float T_final = 0.f, w_z_total = 0.f, w_z2_total = 0.f;
  uint32_t last_contributor = 0;
  float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f;
  float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gp3 = 0.f, gp4 = 0.f, gp5 = 0.f, gp6 = 0.f;
  float gd = 0.f, ga = 0.f, gdd = 0.f;
  float bgdot_c = 0.f, bgdot_p = 0.f;
  if (inside) {
    const float4 st = __ldg(p.img_state + pix_id);
    T_final = st.x;
    w_z_total = st.y;
    w_z2_total = st.z;
    last_contributor = __float_as_uint(st.w);
    gc0 = __ldg(p.dL_dcolor + 0 * HW + pix_id);
    gc1 = __ldg(p.dL_dcolor + 1 * HW + pix_id);
    gc2 = __ldg(p.dL_dcolor + 2 * HW + pix_id);
    gp0 = __ldg(p.dL_dphasor + 0 * HW + pix_id);
    gp1 = __ldg(p.dL_dphasor + 1 * HW + pix_id);
    gp2 = __ldg(p.dL_dphasor + 2 * HW + pix_id);
    gp3 = __ldg(p.dL_dphasor + 3 * HW + pix_id);
    gp4 = __ldg(p.dL_dphasor + 4 * HW + pix_id);
    gp5 = __ldg(p.dL_dphasor + 5 * HW + pix_id);
    gp6 = __ldg(p.dL_dphasor + 6 * HW + pix_id);
    gd = __ldg(p.dL_ddepth + pix_id);
    ga = __ldg(p.dL_dacc + pix_id);
    gdd = __ldg(p.dL_ddd + pix_id);
    float bgv[7];
    if (p.bg_mode == 0) {
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) bgv[ch] = __ldg(p.bg + ch * HW + pix_id);
    } else {
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) bgv[ch] = __ldg(p.bg + ch);
    }
    // backward.cu:850-857 (pixel constants; the reference recomputes them for every pair)
    bgdot_c = bgv[0] * gc0 + bgv[1] * gc1 + bgv[2] * gc2;
    bgdot_p = bgv[0] * gp0 + bgv[1] * gp1 + bgv[2] * gp2 + bgv[3] * gp3 + bgv[4] * gp4 +
              bgv[5] * gp5 + bgv[6] * gp6;
  }

  // furthest contributor of the warp / of the tile
  uint32_t wmax = last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) s.wmax[warp] = wmax;
  __syncthreads();
  uint32_t bmax = 0;
#pragma unroll
  for (int w = 0; w < GFT_BLOCK / 32; ++w) bmax = max(bmax, s.wmax[w]);
  const int n_eff = min(n, (int)bmax);  // Gaussians at tile-list positions >= bmax are skipped by every pixel

  float T = T_final;
  float ar_c0 = 0.f, ar_c1 = 0.f, ar_c2 = 0.f;
  float ar_p0 = 0.f, ar_p1 = 0.f, ar_p2 = 0.f, ar_p3 = 0.f, ar_p4 = 0.f, ar_p5 = 0.f, ar_p6 = 0.f;
  float ar_d = 0.f, ar_a = 0.f, ar_dd = 0.f;
  const float ddelx_dx = 0.5f * (float)p.W;
  const float ddely_dy = 0.5f * (float)p.H;
  const float one_m_Tf = 1.f - T_final;

  const int nb = (n_eff + BATCH - 1) / BATCH;
  for (int b = nb - 1; b >= 0; --b) {
    const int base = b * BATCH;
    const int m = min(BATCH, n_eff - base);
    __syncthreads();
    if ((int)tid < m) {
      const int g = (int)__ldg(p.point_list + range.x + base + tid);
      const float4* r = p.rec + (size_t)g * (GFT_REC_FLOATS / 4);
      s.id[tid] = g;
      s.r0[tid] = __ldg(r + 0);
      s.r1[tid] = __ldg(r + 1);
      s.r2[tid] = __ldg(r + 2);
      s.r3[tid] = __ldg(r + 3);
      s.r4[tid] = __ldg(r + 4);
    }
    __syncthreads();

    // positions >= wmax contribute nothing for this warp
    const int m_w = min(m, (int)wmax - base);
    for (int c = ((m_w - 1) >> 5) << 5; c >= 0 && m_w > 0; c -= 32) {
      const int jj = c + (int)lane;
      bool hit = false;
      if (jj < m_w) {
        const float4 g0 = s.r0[jj];
        hit = !(g0.x + g0.z < patch_x0 || g0.x - g0.z > patch_x1 || g0.y + g0.w < patch_y0 ||
                g0.y - g0.w > patch_y1);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int bsel = 31 - __clz(mask);
        mask &= ~(1u << bsel);
        const int k = c + bsel;
        const float4 g0 = s.r0[k];
        const float4 g1 = s.r1[k];
        const float dx = __fsub_rn(g0.x, pixfx);
        const float dy = __fsub_rn(g0.y, pixfy);
        const float power = pair_power(dx, dy, g1.x, g1.y, g1.z);
        // backward.cu:739-754
        bool contrib = ((uint32_t)(base + k) < last_contributor) && !(power > 0.0f);
        float G = 0.f, alpha = 0.f;
        if (contrib) {
          G = expf(power);
          alpha = fminf(0.99f, __fmul_rn(g1.w, G));
          contrib = !(alpha < 1.0f / 255.0f);
        }
        if (!__any_sync(0xffffffffu, contrib)) continue;

        float v[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) v[i] = 0.f;
        if (contrib) {
          const float4 g2 = s.r2[k];
          const float4 g3 = s.r3[k];
          const float4 g4 = s.r4[k];
          const float om = 1.f - alpha;
          T = T / om;
          const float w = alpha * T;      // dchannel_dcolor, dchannel_ddepth
          const float wp = w * T;         // dchannel_dphasor = alpha * T * T

          // colour (backward.cu:776-790)
          float dLa_c = (g2.x - ar_c0) * gc0 + (g2.y - ar_c1) * gc1 + (g2.z - ar_c2) * gc2;
          dLa_c *= T;
          // phasor (backward.cu:793-804)
          const float two_om = 2.f * om;
          float dLa_p = (g3.x - two_om * ar_p0) * gp0 + (g3.y - two_om * ar_p1) * gp1 +
                        (g3.z - two_om * ar_p2) * gp2 + (g3.w - two_om * ar_p3) * gp3 +
                        (g4.x - two_om * ar_p4) * gp4 + (g4.y - two_om * ar_p5) * gp5 +
                        (g4.z - two_om * ar_p6) * gp6;
          dLa_p *= T * T;
          // depth (backward.cu:807-813)
          const float dLa_d = (g2.w - ar_d) * gd * T;
          // acc (backward.cu:816-818): accum_rec_a is advanced BEFORE use
          const float dLa_a = (1.f - ar_a) * ga * T;
          // depth distortion (backward.cu:825-833), uses (1 - T_final), quirk A.7-5
          const float z = g4.w;
          const float dL_dw = gdd * (z * z * one_m_Tf - 2.0f * z * w_z_total + w_z2_total);
          const float dLa_dd = (dL_dw - ar_dd) * T;
          const float g_ndc = gdd * 2.0f * alpha * T * (z * one_m_Tf - w_z_total);

          // background (backward.cu:850-858); the phasor term is added after the T^2 scaling
          const float bgf = -T_final / om;
          float dL_dalpha = bgf * bgdot_c;
          dLa_p += bgf * bgdot_p;
          dL_dalpha += dLa_c;
          dL_dalpha += dLa_p;
          dL_dalpha += dLa_d;
          dL_dalpha += dLa_a;
          dL_dalpha += dLa_dd;

          // advance the back-to-front recurrences for the next (nearer) Gaussian
          ar_c0 = alpha * g2.x + om * ar_c0;
          ar_c1 = alpha * g2.y + om * ar_c1;
          ar_c2 = alpha * g2.z + om * ar_c2;
          const float om2 = om * om;
          ar_p0 = alpha * g3.x + om2 * ar_p0;
          ar_p1 = alpha * g3.y + om2 * ar_p1;
          ar_p2 = alpha * g3.z + om2 * ar_p2;
          ar_p3 = alpha * g3.w + om2 * ar_p3;
          ar_p4 = alpha * g4.x + om2 * ar_p4;
          ar_p5 = alpha * g4.y + om2 * ar_p5;
          ar_p6 = alpha * g4.z + om2 * ar_p6;
          ar_d = alpha * g2.w + om * ar_d;
          ar_a = alpha + om * ar_a;
          ar_dd = alpha * dL_dw + om * ar_dd;

          // backward.cu:869-886
          const float dL_dG = g1.w * dL_dalpha;
          const float gdx = G * dx, gdy = G * dy;
          const float dG_ddelx = -gdx * g1.x - gdy * g1.y;
          const float dG_ddely = -gdy * g1.z - gdx * g1.y;
          v[0] = dL_dG * dG_ddelx * ddelx_dx;   // mean2D.x
          v[1] = dL_dG * dG_ddely * ddely_dy;   // mean2D.y
          v[2] = -0.5f * gdx * dx * dL_dG;      // conic.x
          v[3] = -0.5f * gdx * dy * dL_dG;      // conic.y
          v[5] = -0.5f * gdy * dy * dL_dG;      // conic.w   (rec 4)
          v[6] = G * dL_dalpha;                 // opacity   (rec 5)
          v[7] = w * gc0;                       // rec 6
          v[8] = w * gc1;                       // rec 7
          v[10] = w * gc2;                      // rec 8
          v[11] = w * gd;                       // rec 9  dist
          v[12] = g_ndc;                        // rec 10 ndc
          v[13] = wp * gp0;                     // rec 11
          v[15] = wp * gp1;                     // rec 12
          v[16] = wp * gp2;                     // rec 13
          v[17] = wp * gp3;                     // rec 14
          v[18] = wp * gp4;                     // rec 15
          v[4] = wp * gp5;                      // rec 16
          v[9] = wp * gp6;                      // rec 17
        }

        // ---- halving butterfly over the warp --------------------------------------------
        float r[10];
        {
          const bool up = (lane & 16u) != 0u;
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            const float send = up ? v[i] : v[i + 10];
            const float keep = up ? v[i + 10] : v[i];
            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        float q[5];
        {
          const bool up = (lane & 8u) != 0u;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float send = up ? r[i] : r[i + 5];
            const float keep = up ? r[i + 5] : r[i];
            q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
          for (int i = 0; i < 5; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], o);
        }
        if ((lane & 7u) == 0u) {
          const uint32_t grp = lane >> 3;
          float* dst = p.grad_rec + (size_t)s.id[k] * GFT_GRAD_FLOATS;
          atomicAdd(reinterpret_cast<float4*>(dst) + grp, make_float4(q[0], q[1], q[2], q[3]));
          if (grp < 2u) atomicAdd(dst + 16 + grp, q[4]);
        }
      }
    }
  }
}

void launch_blend_bwd(const BlendBwdParams& p, cudaStream_t stream) {
  const int tiles = p.grid_x * p.grid_y;
  if (tiles <= 0) return;
  blend_bwd_kernel<<<tiles, GFT_BLOCK, 0, stream>>>(p);
  note_launches(1);
}

}  // namespace gft
