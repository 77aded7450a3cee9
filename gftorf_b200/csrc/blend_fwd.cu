// blend_fwd.cu — per-tile front-to-back blend of RGB, 7-channel ToF phasor/quads, depth,
// accumulation, depth distortion and first-hit distribution, one pass, for sm_100a.
//
// Replaces renderCUDA<3,7,2,1> (cuda_rasterizer/forward.cu:424-676).
//
// What is different from the reference kernel, and why the results are still the same:
//   - every per-Gaussian quantity (xy, conic+opacity, rgb, 7 phasor channels, dist, ndc) is
//     staged in shared memory from ONE 80-byte record per Gaussian, gathered with cp.async into a
//     double buffer so the next batch's gather overlaps the blending of the current one; the reference stages 7 floats
//     and re-gathers 12 more from global memory for every contributing (pixel, Gaussian) pair
//     (forward.cu:551-572).
//   - a 16x16 tile is covered by 8 warps of 8x4 pixels.  Each warp tests, 32 Gaussians at a time,
//     whether a Gaussian's conservative alpha>=1/255 box reaches its 8x4 patch and only walks the
//     survivors.  A culled pair would have hit `continue` in the reference before touching any
//     state (forward.cu:528-537), so skipping it is exact.
//   - the alpha chain (power, expf, min, the two thresholds, T update) is evaluated with the
//     exact operation order of the reference binary, so `done`, n_contrib and pixels are
//     bit-identical.
//   - `pixels` (one global float atomic per contributing pair in the reference, forward.cu:629)
//     is counted with a warp ballot, accumulated per tile in shared memory and flushed with one
//     atomic per (tile, Gaussian); the counts are integers < 2^24, so any order is exact.
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace gft {

namespace {

// Warp-autonomous blend: a 16x16 tile is covered by 8 warps of 8x4 pixels that never synchronise
// with each other.  Every warp gathers the tile's Gaussian list for itself, 32 Gaussians (one per
// lane) at a time, with cp.async into a warp-private ring of WSTAGES slots of 80-byte records, and
// walks it at its own pace.  (A block-synchronous version — one shared double buffer, one
// __syncthreads per batch of 256 — had 25 % of its warp samples parked at the barrier, waiting for
// the busiest warp of the tile; it was 4 % slower at 640x480 and 2 % at 1080p and is gone.)  The
// price is that the 8 warps each read the records (L1/L2 hits after the first) and flush their
// own `pixels` counts (one atomic per (warp, Gaussian) with a contribution).
//
// One launch covers the tiles of ALL views of the batch: block b works on global tile b, which
// belongs to the view v with tile_base_v <= b < tile_base_v + T_v.
constexpr int WSTAGES = 2;
struct WarpStage {
  float4 r0[32], r1[32], r2[32], r3[32], r4[32];
  int id[32];
};

__device__ __forceinline__ void cp_async16_ca(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace

// HALF: the warp's 8x4 patch is walked as two independent 4x4 halves (lanes 0-15 / 16-31).  Each
// half has its own list of candidates that can reach it and the two lists are walked side by side,
// so one instruction stream blends two (half-patch, Gaussian) pairs at once: a 4x4 half is hit by
// fewer Gaussians than the 8x4 patch and a larger share of its 16 lanes contributes.  Per pixel
// the candidates still arrive in list order with the same arithmetic: results are unchanged.
template <int MINB, bool HALF>
__global__ void __launch_bounds__(GFT_BLOCK, MINB)
blend_fwd_warp_kernel(const __grid_constant__ BlendFwdParams p) {
  extern __shared__ __align__(16) unsigned char fwd_smem_raw[];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WarpStage* ring = reinterpret_cast<WarpStage*>(fwd_smem_raw) + warp * WSTAGES;
  // the view this tile belongs to (uniform per block)
  const uint32_t tile_g = p.order ? __ldg(p.order + blockIdx.x) : blockIdx.x;   // longest lists first
  int vi = 0;
  for (int v = 1; v < p.nviews; ++v) vi = ((int)tile_g >= p.views[v].tile_base) ? v : vi;
  const BlendViewFwd& vw = p.views[vi];
  const uint32_t tile = tile_g - (uint32_t)vw.tile_base;
  const uint32_t tile_x = tile % (uint32_t)vw.grid_x, tile_y = tile / (uint32_t)vw.grid_x;
  const uint32_t px0 = tile_x * GFT_TILE_X + (warp & 1u) * 8u;
  const uint32_t py0 = tile_y * GFT_TILE_Y + (warp >> 1) * 4u;
  const uint32_t half = lane >> 4;                      // HALF: 0 = left 4x4, 1 = right 4x4
  const uint32_t pix_x = HALF ? px0 + half * 4u + (lane & 3u) : px0 + (lane & 7u);
  const uint32_t pix_y = HALF ? py0 + ((lane & 15u) >> 2) : py0 + (lane >> 3);
  const int W = vw.W, H = vw.H;
  const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
  const uint32_t pix_id = (uint32_t)W * pix_y + pix_x;
  const float pixfx = (float)pix_x, pixfy = (float)pix_y;
  const float patch_x0 = (float)px0, patch_x1 = (float)(px0 + 7u);
  const float patch_y0 = (float)py0, patch_y1 = (float)(py0 + 3u);
  const float4* __restrict__ recs = vw.rec;
  float* __restrict__ pixels = vw.pixels;

  const uint2 range = p.ranges[tile_g];
  const int n = (int)(range.y - range.x);
  const int nb = (n + 31) >> 5;

  bool done = !inside;
  float T = 1.0f;
  uint32_t last_contributor = 0;
  // accumulators in pairs for the packed FFMA2: (C0,C1) (C2,D) use weight w, (P0,P1) (P2,P3)
  // (P4,P5) use weight w*T — the pairs are the halves of the 128-bit record loads
  float2 C01 = make_float2(0.f, 0.f), C2D = C01, P01 = C01, P23 = C01, P45 = C01;
  float P6 = 0.f;
  float A = 0.f, DD = 0.f, DD_D = 0.f, DD_D2 = 0.f;
  float WD0 = 0.f, WD1 = 0.f, WD2 = 0.f;
  bool first_hit = true;

  auto issue = [&](int b) {
    if (b < nb) {
      const int j = b * 32 + (int)lane;
      if (j < n) {
        const int g = (int)__ldg(p.point_list + range.x + j);
        const float4* r = recs + (size_t)g * (GFT_REC_FLOATS / 4);
        WarpStage& d = ring[b % WSTAGES];
        d.id[lane] = g;
        cp_async16_ca(&d.r0[lane], r + 0);
        cp_async16_ca(&d.r1[lane], r + 1);
        cp_async16_ca(&d.r2[lane], r + 2);
        cp_async16_ca(&d.r3[lane], r + 3);
        cp_async16_ca(&d.r4[lane], r + 4);
      }
    }
    cp_async_commit();     // one group per batch slot, empty or not, so the group count stays uniform
  };

  if (!__all_sync(0xffffffffu, done)) {
#pragma unroll
    for (int b = 0; b < WSTAGES - 1; ++b) issue(b);
    for (int b = 0; b < nb; ++b) {
      issue(b + WSTAGES - 1);
      cp_async_wait_group<WSTAGES - 1>();   // this lane's copies of batch b have landed ...
      __syncwarp();                         // ... and so have the other lanes'
      const WarpStage& s = ring[b % WSTAGES];
      const int base = b * 32;
      const int m = min(32, n - base);
      int mycnt = 0;
      // alpha of one pair (forward.cu:524-537), independent of the pixel's running state, so two
      // Gaussians are evaluated side by side to overlap their expf latency chains
      auto eval_alpha = [&](int k, float& alpha) -> bool {
        const float4 g0 = s.r0[k];
        const float4 g1 = s.r1[k];
        const float dx = __fsub_rn(g0.x, pixfx);
        const float dy = __fsub_rn(g0.y, pixfy);
        const float power = pair_power(dx, dy, g1.x, g1.y, g1.z);
        const bool neg = !(power > 0.0f);
        alpha = fminf(0.99f, __fmul_rn(g1.w, expf(neg ? power : 0.0f)));
        return neg && !(alpha < 1.0f / 255.0f);
      };
      // One (pixel, Gaussian) pair, after its alpha has been evaluated.  Sequential per pixel: the
      // transmittance test and the accumulators depend on every earlier Gaussian.  kL / kR: the
      // candidates of the left / right half in this step (the same candidate without HALF); the
      // lane that staged a candidate collects its pixel count.
      auto apply = [&](int k, int kL, int kR, float alpha, bool pass) {
        bool contrib = !done && pass;
        float test_T = 0.f;
        if (contrib) {
          test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
          if (test_T < 0.0001f) { done = true; contrib = false; }   // forward.cu:538-543
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, contrib);
        if (HALF) {
          if ((int)lane == kL) mycnt += __popc(bal & 0xffffu);
          if ((int)lane == kR) mycnt += __popc(bal >> 16);
        } else {
          if ((int)lane == k) mycnt = __popc(bal);
        }
        if (contrib) {
          const float4 g2 = s.r2[k];
          const float4 g3 = s.r3[k];
          const float4 g4 = s.r4[k];
          const float w = __fmul_rn(T, alpha);
          const float wp = __fmul_rn(T, w);
          const float2 w2 = make_float2(w, w), wp2 = make_float2(wp, wp);
          C01 = fma2(w2, make_float2(g2.x, g2.y), C01);
          C2D = fma2(w2, make_float2(g2.z, g2.w), C2D);    // .y is the depth accumulator D
          P01 = fma2(wp2, make_float2(g3.x, g3.y), P01);
          P23 = fma2(wp2, make_float2(g3.z, g3.w), P23);
          P45 = fma2(wp2, make_float2(g4.x, g4.y), P45);
          P6 = __fmaf_rn(wp, g4.z, P6);
          if (first_hit) { WD0 = alpha; WD1 = g2.w; WD2 = g3.z; first_hit = false; }   // forward.cu:561-567
          // depth distortion, forward.cu:572-578 in the reference's evaluation order
          const float z = g4.w;
          const float z2 = __fmul_rn(z, z);
          const float t1 = __fmul_rn(DD_D, __fadd_rn(z, z));
          float t2 = __fmaf_rn(A, z2, -t1);
          const float wz = __fmul_rn(w, z);
          t2 = __fadd_rn(DD_D2, t2);
          DD_D = __fadd_rn(DD_D, wz);
          DD_D2 = __fmaf_rn(z, wz, DD_D2);
          DD = __fmaf_rn(w, t2, DD);
          A = __fadd_rn(A, w);
          T = test_T;
          last_contributor = (uint32_t)(base + k + 1);
        }
      };
      if (HALF) {
        bool hitL = false, hitR = false;
        if ((int)lane < m) {
          const float4 g0 = s.r0[lane];
          // written so that NaN extents mean "not culled"
          const bool ymiss = g0.y + g0.w < patch_y0 || g0.y - g0.w > patch_y1;
          hitL = !(ymiss || g0.x + g0.z < patch_x0 || g0.x - g0.z > patch_x0 + 3.f);
          hitR = !(ymiss || g0.x + g0.z < patch_x0 + 4.f || g0.x - g0.z > patch_x1);
        }
        uint32_t mL = __ballot_sync(0xffffffffu, hitL), mR = __ballot_sync(0xffffffffu, hitR);
        const uint32_t dn = __ballot_sync(0xffffffffu, done);
        if ((dn & 0xffffu) == 0xffffu) mL = 0u;        // a half whose pixels are all done stops walking
        if ((dn >> 16) == 0xffffu) mR = 0u;
        while (mL | mR) {
          const int kL1 = mL ? __ffs(mL) - 1 : -1;
          if (mL) mL &= mL - 1;
          const int kR1 = mR ? __ffs(mR) - 1 : -1;
          if (mR) mR &= mR - 1;
          const int kL2 = mL ? __ffs(mL) - 1 : -1;
          if (mL) mL &= mL - 1;
          const int kR2 = mR ? __ffs(mR) - 1 : -1;
          if (mR) mR &= mR - 1;
          const int k1 = half ? kR1 : kL1, k2 = half ? kR2 : kL2;
          float alpha1 = 0.f, alpha2 = 0.f;
          const bool pass1 = k1 >= 0 && eval_alpha(k1, alpha1);
          const bool pass2 = k2 >= 0 && eval_alpha(k2, alpha2);
          apply(max(k1, 0), kL1, kR1, alpha1, pass1);
          if (kL2 >= 0 || kR2 >= 0) apply(max(k2, 0), kL2, kR2, alpha2, pass2);
        }
      } else {
        bool hit = false;
        if ((int)lane < m) {
          const float4 g0 = s.r0[lane];
          // written so that NaN extents mean "not culled"
          hit = !(g0.x + g0.z < patch_x0 || g0.x - g0.z > patch_x1 || g0.y + g0.w < patch_y0 ||
                  g0.y - g0.w > patch_y1);
        }
        uint32_t mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int b1 = __ffs(mask) - 1;
          mask &= mask - 1;
          const bool two = mask != 0u;
          const int b2 = two ? (__ffs(mask) - 1) : b1;
          if (two) mask &= mask - 1;
          float alpha1, alpha2;
          const bool pass1 = eval_alpha(b1, alpha1);
          const bool pass2 = eval_alpha(b2, alpha2);
          apply(b1, b1, b1, alpha1, pass1);
          if (two) apply(b2, b2, b2, alpha2, pass2);
        }
      }
      // `pixels` (forward.cu:629: one float atomic per contributing pair) counted by ballot and
      // flushed once per (warp, Gaussian); integer-valued floats < 2^24, exact in any order
      if (mycnt) atomicAdd(pixels + s.id[lane], (float)mycnt);
      const bool all_done = __all_sync(0xffffffffu, done);   // also orders this batch's reads before the next issue
      if (all_done) break;
    }
    cp_async_wait_group<0>();
  }

  if (inside) {
    const size_t HW = (size_t)H * (size_t)W;
    p.img_state[vw.pix_base + pix_id] = make_float4(T, DD_D, DD_D2, __uint_as_float(last_contributor));
    float bgv[7];
    if (vw.bg_mode == 0) {
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) bgv[ch] = __ldg(vw.bg + ch * HW + pix_id);
    } else {
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) bgv[ch] = __ldg(vw.bg + ch);
    }
    // colour reads bg planes 0..2, phasor planes 0..6, both weighted by T (forward.cu:644,649)
    vw.out_color[0 * HW + pix_id] = __fmaf_rn(T, bgv[0], C01.x);
    vw.out_color[1 * HW + pix_id] = __fmaf_rn(T, bgv[1], C01.y);
    vw.out_color[2 * HW + pix_id] = __fmaf_rn(T, bgv[2], C2D.x);
    vw.out_phasor[0 * HW + pix_id] = __fmaf_rn(T, bgv[0], P01.x);
    vw.out_phasor[1 * HW + pix_id] = __fmaf_rn(T, bgv[1], P01.y);
    vw.out_phasor[2 * HW + pix_id] = __fmaf_rn(T, bgv[2], P23.x);
    vw.out_phasor[3 * HW + pix_id] = __fmaf_rn(T, bgv[3], P23.y);
    vw.out_phasor[4 * HW + pix_id] = __fmaf_rn(T, bgv[4], P45.x);
    vw.out_phasor[5 * HW + pix_id] = __fmaf_rn(T, bgv[5], P45.y);
    vw.out_phasor[6 * HW + pix_id] = __fmaf_rn(T, bgv[6], P6);
    vw.out_depth[pix_id] = C2D.y;
    vw.out_acc[pix_id] = A;
    vw.out_depth_distortion[pix_id] = DD;
    vw.out_distribution[0 * HW + pix_id] = WD0;
    vw.out_distribution[1 * HW + pix_id] = WD1;
    vw.out_distribution[2 * HW + pix_id] = WD2;
    // outputs the reference allocates but never writes (forward.cu:655-667): defined as 0
    if (vw.out_normal) {
      vw.out_normal[0 * HW + pix_id] = 0.f;
      vw.out_normal[1 * HW + pix_id] = 0.f;
      vw.out_normal[2 * HW + pix_id] = 0.f;
    }
    if (vw.out_entropy) vw.out_entropy[pix_id] = 0.f;
    if (vw.out_amp_distortion) vw.out_amp_distortion[pix_id] = 0.f;
  }
}

void launch_blend_fwd(const BlendFwdParams& p, cudaStream_t stream) {
  if (p.T_total <= 0) return;
  const int smem = (GFT_BLOCK / 32) * WSTAGES * (int)sizeof(WarpStage);
  // 4 resident blocks per SM: 3 (67 registers) is 4-7 % slower, 5 (48 registers, 8 B spilled)
  // 10-15 % slower on B200.  Option blend_half selects the two-halves walk.
  static unsigned long long ok_full = 0, ok_half = 0;
  if (option(OPT_BLEND_HALF) != 0) {
    ensure_dynamic_smem(blend_fwd_warp_kernel<4, true>, smem, &ok_half);
    blend_fwd_warp_kernel<4, true><<<p.T_total, GFT_BLOCK, smem, stream>>>(p);
  } else {
    ensure_dynamic_smem(blend_fwd_warp_kernel<4, false>, smem, &ok_full);
    blend_fwd_warp_kernel<4, false><<<p.T_total, GFT_BLOCK, smem, stream>>>(p);
  }
  note_launches(1);
}

}  // namespace gft
