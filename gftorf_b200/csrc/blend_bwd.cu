// blend_bwd.cu — back-to-front replay of the per-tile blend and its gradients, for sm_100a.
//
// Replaces renderCUDA<3,7> of the backward pass (cuda_rasterizer/backward.cu:609-889).
//
// The reference keeps 13 per-channel back-to-front recurrences per pixel (accum_rec[3],
// accum_rec_p[7], accum_rec_d, accum_rec_a, accum_rec_dd; backward.cu:673-683,780-833) and issues
// 18 scalar global atomicAdds for every contributing (pixel, Gaussian) pair
// (:788,802,812,832,876-877,881-883,886).  Here:
//
//   * Scalar recurrences.  dL/dalpha only ever needs the recurrences CONTRACTED with the pixel's
//     incoming gradients, and those contractions obey the same recurrences:
//        weight alpha*T   family (colour, depth, acc, distortion):  X <- alpha*x + (1-alpha)*X,
//            x = sum_c c*g_c + dist*g_d + dL_dw + g_a,      dL/dalpha += (x - X) * T
//        weight alpha*T^2 family (7 phasor channels):            B <- alpha*pi + (1-alpha)^2*B,
//            pi = sum_ch p_ch*g_ch,                          dL/dalpha += (pi - 2(1-alpha)B) * T^2
//     Two scalar recurrences replace thirteen; same mathematics, a different (shorter) rounding path.
//   * Per-Gaussian partials are sums over pixels of (per-pair scalar) x (per-pixel constant), and
//     per-Gaussian factors are pulled out of the sums: the kernel accumulates
//        S_x=sum h*dx  S_y=sum h*dy  S_xx=sum h*dx^2  S_xy=sum h*dx*dy  S_yy=sum h*dy^2   (h = o*G*dL/dalpha)
//        sum G*dL/dalpha | sum w*g_c[3] | sum w*g_d | sum w*(2z*k2-k1) | sum wp*(gA,gB,gC,gS)
//     where (gA,gB,gC,gS) are the four linear combinations of the 7 phasor pixel gradients that the
//     phasor backward consumes (backward.cu:551-577).  15 values instead of 18.
//   * Warp reduction of the 15 partials: a transpose through a warp-private shared buffer — every
//     lane stores its 15 values down one column (15 STS, conflict-free), lane j sums half a row
//     with four 128-bit loads + packed adds, one shuffle joins the halves: ~38 instructions (the
//     halving shuffle butterfly it replaced cost ~64 and was 30 % of the kernel's instructions).
//     Value i ends in lanes 2i, 2i+1 and the even lanes issue ONE reduction instruction onto the
//     Gaussian's 64-byte record.
//   * No divergent region in the replay (PRED): a pair that does not contribute is replayed with
//     alpha = G = 0, for which every partial is exactly 0 and the recurrences leave the pixel's
//     state untouched (T/1, 0*x + 1*X), so the `if (contributes)` of the reference
//     (backward.cu:744-754) needs no branch, no re-convergence barrier and no zero fill of the 15
//     partials.
//   * a 16x16 tile is walked by 8 warps of 8x4 pixels, 32 Gaussians at a time; a warp skips
//     Gaussians whose conservative alpha>=1/255 box misses its patch (exact, see blend_fwd.cu);
//     batches behind the tile's furthest last-contributor are never loaded; the next batch is
//     gathered with cp.async into the other half of a double buffer while this one is processed.
//
// Gradient record [P][16] (private; consumed by preprocess_bwd.cu):
//   0 S_x  1 S_y  2 S_xx  3 S_xy  4 S_yy  5 opacity  6 col.r  7 col.g  8 col.b  9 dist  10 ndc
//   11 phA = dR+dq1-dq2   12 phB = dI+dq3-dq4   13 phC = dA   14 phS = dq1+dq2+dq3+dq4   15 unused
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace gft {

namespace {

constexpr int RED_STRIDE = 36;
constexpr int RED_FLOATS = 16 * RED_STRIDE;
// HALF: 32 rows (15 values x 2 halves) of 16 lanes, 20-float stride, the right half's rows shifted by
// 16 floats so the two halves' stores hit disjoint banks
constexpr int REDH_STRIDE = 20;
constexpr int REDH_FLOATS = 32 * REDH_STRIDE + 16;

template <int BATCH>
struct BwdBufT {
  float4 r0[BATCH];  // x y ex ey
  float4 r1[BATCH];  // conA conB conC opacity
  float4 r2[BATCH];  // r g b dist
  float4 r3[BATCH];  // ph0..ph3
  float4 r4[BATCH];  // ph4 ph5 ph6 ndc
  int id[BATCH];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 1 / (1 - alpha) for 1 - alpha in [0.01, 1]: the hardware approximation refined by one Newton step
// (3 instructions, error below 1 ulp) instead of the IEEE division's range check, branch and slow
// path (10 instructions in the replay's dependency chain).
__device__ __forceinline__ float recip_om(float om) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(om));
  return fmaf(fmaf(-om, r, 1.0f), r, r);
}

// ---- ring mode: 16 slots of 32 Gaussians, one mbarrier pair per slot -----------------------------
constexpr int RING_SLOTS = 16;
constexpr int RING_LEAD = 8;     // a chunk is requested 8 chunks before its use; warp w requests chunks w, w+8, ...
struct RingSlot {
  float4 r0[32], r1[32], r2[32], r3[32], r4[32];
  int id[32];
};
static_assert(RING_SLOTS * sizeof(RingSlot) == 2 * sizeof(BwdBufT<GFT_BLOCK>), "ring and double buffer share the staging area");

__device__ __forceinline__ void cp_async4_ca(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared.b64 t, [%0]; }" ::"r"(a) : "memory");
}
// arrive-on triggered when all cp.async operations this thread issued before have landed
__device__ __forceinline__ void mbar_arrive_on_copies(unsigned long long* bar) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) break;
    if (clock64() - t0 > 8000000000ll) __trap();   // seconds: a protocol error must not hang the GPU
  }
}

}  // namespace

// 8 warps = the eight 8x4 patches of one 16x16 tile (see blend_fwd.cu), BATCH = 256 Gaussians
// staged per step (one per thread) into a double buffer.  One launch covers the tiles of all
// views of the batch (global tile index = blockIdx.x).
// RING: instead of one double buffer of 256 Gaussians filled by the whole block between two
// __syncthreads (ncu: 18 % of the warp samples wait there for the busiest warp of the tile), the
// staging area is a ring of 16 slots x 32 Gaussians with a full / empty mbarrier pair per slot.
// Warp w requests chunks w, w+8, ... eight chunks ahead of their use (cp.async, the slot's `full`
// barrier fires when the 32 lanes' copies have landed) and every warp releases a slot when it is
// done with it, so a warp only ever waits for a chunk that has not arrived or for the slowest
// warp falling more than 8 chunks behind — never for the block.  Every wait depends on strictly
// earlier chunks only, so the protocol cannot deadlock.
// HALF (with RING and PRED): the warp's 8x4 patch is replayed as two independent 4x4 halves, each
// with its own candidate list, side by side in one instruction stream (see blend_fwd.cu); the two
// halves' 15 partials are reduced by ONE transpose (32 rows of 16 lanes, lane j sums row j, no
// shuffle) and leave with one reduction instruction for both Gaussians.
template <int MINB, bool PRED, bool RING, bool HALF = false>
__global__ void __launch_bounds__(GFT_BLOCK, MINB)
blend_bwd_kernel(const __grid_constant__ BlendBwdParams p) {
  constexpr int WARPS = GFT_BLOCK / 32;
  constexpr int BATCH = GFT_BLOCK;
  using BwdBuf = BwdBufT<BATCH>;
  extern __shared__ __align__(16) unsigned char bwd_smem_raw[];
  BwdBuf* buf = reinterpret_cast<BwdBuf*>(bwd_smem_raw);
  RingSlot* ring = reinterpret_cast<RingSlot*>(bwd_smem_raw);
  __shared__ uint32_t s_wmax[WARPS];
  __shared__ __align__(8) unsigned long long s_full[RING_SLOTS], s_empty[RING_SLOTS];
  // per-warp transpose buffer, 16 value rows x 36 floats (32 lanes + 4 pad)
  float* red = reinterpret_cast<float*>(bwd_smem_raw + 2 * sizeof(BwdBuf)) + (threadIdx.x >> 5) * (HALF ? REDH_FLOATS : RED_FLOATS);

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile_g = p.order ? __ldg(p.order + blockIdx.x) : blockIdx.x;   // longest lists first
  int vi = 0;
  for (int v = 1; v < p.nviews; ++v) vi = ((int)tile_g >= p.views[v].tile_base) ? v : vi;
  const BlendViewBwd& vw = p.views[vi];
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
  const size_t HW = (size_t)H * (size_t)W;
  const float4* __restrict__ recs = vw.rec;
  float* __restrict__ grad_rec = vw.grad_rec;

  const uint2 range = p.ranges[tile_g];
  const int n = (int)(range.y - range.x);

  // ---- per-pixel forward state, incoming gradients and the constants derived from them ----
  float T_final = 0.f;
  uint32_t last_contributor = 0;
  float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f;
  float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gp3 = 0.f, gp4 = 0.f, gp5 = 0.f, gp6 = 0.f;
  float gd = 0.f, k0 = 0.f, k1 = 0.f, k2 = 0.f, bgdot = 0.f;
  if (inside) {
    const float4 st = __ldg(p.img_state + vw.pix_base + pix_id);
    T_final = st.x;
    const float w_z_total = st.y, w_z2_total = st.z;
    last_contributor = __float_as_uint(st.w);
    gc0 = __ldg(vw.dL_dcolor + 0 * HW + pix_id);
    gc1 = __ldg(vw.dL_dcolor + 1 * HW + pix_id);
    gc2 = __ldg(vw.dL_dcolor + 2 * HW + pix_id);
    gp0 = __ldg(vw.dL_dphasor + 0 * HW + pix_id);
    gp1 = __ldg(vw.dL_dphasor + 1 * HW + pix_id);
    gp2 = __ldg(vw.dL_dphasor + 2 * HW + pix_id);
    gp3 = __ldg(vw.dL_dphasor + 3 * HW + pix_id);
    gp4 = __ldg(vw.dL_dphasor + 4 * HW + pix_id);
    gp5 = __ldg(vw.dL_dphasor + 5 * HW + pix_id);
    gp6 = __ldg(vw.dL_dphasor + 6 * HW + pix_id);
    gd = __ldg(vw.dL_ddepth + pix_id);
    const float ga = __ldg(vw.dL_dacc + pix_id);
    const float gdd = __ldg(vw.dL_ddd + pix_id);
    // dL_dw + g_a = z^2*k2 - z*k1 + k0  (backward.cu:828 with the (1 - T_final) of quirk A.7-5)
    k2 = gdd * (1.f - T_final);
    k1 = 2.0f * gdd * w_z_total;
    k0 = gdd * w_z2_total + ga;
    float bgv[7];
    if (vw.bg_mode == 0) {
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) bgv[ch] = __ldg(vw.bg + ch * HW + pix_id);
    } else {
#pragma unroll
      for (int ch = 0; ch < 7; ++ch) bgv[ch] = __ldg(vw.bg + ch);
    }
    // backward.cu:850-858: both background terms enter dL/dalpha with the factor -T_final/(1-alpha)
    bgdot = bgv[0] * gc0 + bgv[1] * gc1 + bgv[2] * gc2 +
            (bgv[0] * gp0 + bgv[1] * gp1 + bgv[2] * gp2 + bgv[3] * gp3 + bgv[4] * gp4 +
             bgv[5] * gp5 + bgv[6] * gp6);
  }
  // the four combinations of phasor pixel gradients the phasor backward needs
  const float gA = gp0 + gp3 - gp4;
  const float gB = gp1 + gp5 - gp6;
  const float gS = (gp3 + gp4) + (gp5 + gp6);

  // furthest contributor of the warp / of the tile
  uint32_t wmax = last_contributor;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  const uint32_t hmax = wmax;                            // furthest contributor of this lane's half
  wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, 16));
  if (lane == 0) s_wmax[warp] = wmax;
  if (RING && tid < RING_SLOTS) {
    mbar_init(&s_full[tid], 32);       // one arrive per lane of the requesting warp, when its copies land
    mbar_init(&s_empty[tid], WARPS);   // one arrive per warp of the block, when it is done with the slot
  }
  __syncthreads();
  uint32_t bmax = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) bmax = max(bmax, s_wmax[w]);
  const int n_eff = min(n, (int)bmax);  // positions >= bmax are skipped by every pixel of the tile

  float T = T_final;
  float X = 0.f;     // contracted alpha*T-family recurrence
  float Bp = 0.f;    // contracted alpha*T^2-family recurrence

  const int nb = (n_eff + BATCH - 1) / BATCH;

  auto stage = [&](int b, int which) {
    const int base = b * BATCH;
    const int m = min(BATCH, n_eff - base);
    if ((int)tid < m) {
      const int g = (int)__ldg(p.point_list + range.x + base + tid);
      const float4* r = recs + (size_t)g * (GFT_REC_FLOATS / 4);
      BwdBuf& d = buf[which];
      d.id[tid] = g;
      cp_async16(&d.r0[tid], r + 0);
      cp_async16(&d.r1[tid], r + 1);
      cp_async16(&d.r2[tid], r + 2);
      cp_async16(&d.r3[tid], r + 3);
      cp_async16(&d.r4[tid], r + 4);
    }
    cp_async_commit();
  };

  // backward.cu:739-754 — the same alpha chain as the forward, so the same pairs are replayed
  auto eval_pair = [&](const auto& s, int base, int k, float& G, float& alpha, float& dx,
                       float& dy) -> bool {
    const float4 g0 = s.r0[k];
    const float4 g1 = s.r1[k];
    dx = __fsub_rn(g0.x, pixfx);
    dy = __fsub_rn(g0.y, pixfy);
    const float power = pair_power(dx, dy, g1.x, g1.y, g1.z);
    const bool neg = !(power > 0.0f);
    G = expf(neg ? power : 0.0f);
    alpha = fminf(0.99f, __fmul_rn(g1.w, G));
    return ((uint32_t)(base + k) < last_contributor) && neg && !(alpha < 1.0f / 255.0f);
  };

  auto replay = [&](const auto& s, int k, bool contrib, float G, float alpha, float dx, float dy) {
    if (!__any_sync(0xffffffffu, contrib)) return;
    float v[16];
    if (PRED) {
      // a pair that does not contribute is replayed with alpha = G = 0: every partial below is
      // then exactly 0 and T, X, Bp keep their values (T * 1, 0 * x + 1 * X)
      alpha = contrib ? alpha : 0.f;
      G = contrib ? G : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    if (PRED || contrib) {
      const float4 g1 = s.r1[k];
      const float4 g2 = s.r2[k];
      const float4 g3 = s.r3[k];
      const float4 g4 = s.r4[k];
      const float om = 1.f - alpha;
      const float inv_om = recip_om(om);
      T = T * inv_om;                 // backward.cu:756
      const float w = alpha * T;      // weight of the alpha*T family
      const float wp = w * T;         // weight of the phasor family, alpha*T*T

      const float z = g4.w;
      // the two dot products with packed fp32x2 FMAs (pairs = halves of the 128-bit record loads)
      const float2 kd = fma2(make_float2(g2.z, g2.w), make_float2(gc2, gd),
                             mul2(make_float2(g2.x, g2.y), make_float2(gc0, gc1)));
      const float x_tot = (kd.x + kd.y) + (z * (z * k2 - k1) + k0);
      const float2 pd = fma2(make_float2(g4.x, g4.y), make_float2(gp4, gp5),
                             fma2(make_float2(g3.z, g3.w), make_float2(gp2, gp3),
                                  mul2(make_float2(g3.x, g3.y), make_float2(gp0, gp1))));
      const float pi = (pd.x + pd.y) + g4.z * gp6;
      const float dL_dalpha = (x_tot - X) * T + (pi - 2.f * om * Bp) * (T * T) -
                              (T_final * inv_om) * bgdot;
      // X <- alpha x + om X ;  Bp <- alpha pi + om^2 Bp
      const float2 rec = fma2(make_float2(alpha, alpha), make_float2(x_tot, pi),
                              mul2(make_float2(om, om * om), make_float2(X, Bp)));
      X = rec.x;
      Bp = rec.y;

      const float h = g1.w * dL_dalpha * G;   // dL_dG * G
      const float2 dxy = make_float2(dx, dy);
      const float2 hxy = mul2(make_float2(h, h), dxy);            // h dx, h dy
      const float2 hxx = mul2(make_float2(hxy.x, hxy.x), dxy);    // h dx dx, h dx dy
      v[0] = hxy.x;
      v[1] = hxy.y;
      v[2] = hxx.x;
      v[3] = hxx.y;
      v[4] = hxy.y * dy;
      v[5] = G * dL_dalpha;
      const float2 w2 = make_float2(w, w), wp2 = make_float2(wp, wp);
      const float2 v67 = mul2(w2, make_float2(gc0, gc1)), v89 = mul2(w2, make_float2(gc2, gd));
      const float2 vAB = mul2(wp2, make_float2(gA, gB)), vCS = mul2(wp2, make_float2(gp2, gS));
      v[6] = v67.x; v[7] = v67.y; v[8] = v89.x; v[9] = v89.y;
      v[10] = w * (2.f * z * k2 - k1);
      v[11] = vAB.x; v[12] = vAB.y; v[13] = vCS.x; v[14] = vCS.y;
    }

    // Transpose through shared memory: lane l stores its 15 partials down column l (stride 36:
    // conflict-free), then lane j sums one half (16 lanes) of row j>>1 with four 128-bit loads
    // (the 8 lanes of a load phase hit 8 different bank groups) and one shuffle joins the halves.
#pragma unroll
    for (int i = 0; i < 15; ++i) red[i * RED_STRIDE + (int)lane] = v[i];
    __syncwarp();
    const float4* rp = reinterpret_cast<const float4*>(red + (lane >> 1) * RED_STRIDE + (lane & 1u) * 16u);
    const float4 q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3];
    // 16 -> 1 with packed adds: 7 FADD2 + 1 FADD instead of 15 FADD
    const float2 s0 = add2(make_float2(q0.x, q0.y), make_float2(q0.z, q0.w));
    const float2 s1 = add2(make_float2(q1.x, q1.y), make_float2(q1.z, q1.w));
    const float2 s2 = add2(make_float2(q2.x, q2.y), make_float2(q2.z, q2.w));
    const float2 s3 = add2(make_float2(q3.x, q3.y), make_float2(q3.z, q3.w));
    const float2 st = add2(add2(s0, s1), add2(s2, s3));
    float sum = st.x + st.y;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    __syncwarp();
    if ((lane & 1u) == 0u && lane != 30u)
      atomicAdd(grad_rec + (size_t)s.id[k] * GFT_GRAD_FLOATS + (lane >> 1), sum);
  };

  // one chunk of <= 32 candidates already in shared memory: cull against the warp's patch, replay
  // the survivors back to front, two at a time (their alpha chains are independent of the pixel's
  // running state and overlap; the replay itself stays sequential)
  auto walk = [&](const auto& s, int base, int c, int m_w) {
    const int jj = c + (int)lane;
    bool hit = false;
    if (jj < m_w) {
      const float4 g0 = s.r0[jj];
      hit = !(g0.x + g0.z < patch_x0 || g0.x - g0.z > patch_x1 || g0.y + g0.w < patch_y0 ||
              g0.y - g0.w > patch_y1);
    }
    uint32_t mask = __ballot_sync(0xffffffffu, hit);
    while (mask) {
      const int s1 = 31 - __clz(mask);
      mask &= ~(1u << s1);
      const bool two = mask != 0u;
      const int s2 = two ? (31 - __clz(mask)) : s1;
      if (two) mask &= ~(1u << s2);
      float G1, G2, al1, al2, dx1, dy1, dx2, dy2;
      const bool c1 = eval_pair(s, base, c + s1, G1, al1, dx1, dy1);
      const bool c2 = eval_pair(s, base, c + s2, G2, al2, dx2, dy2);
      replay(s, c + s1, c1, G1, al1, dx1, dy1);
      if (two) replay(s, c + s2, c2, G2, al2, dx2, dy2);
    }
  };

  // ---- HALF: two candidates per step, one per 4x4 half ------------------------------------------
  auto replay_half = [&](const auto& s, int k, bool contrib, float G, float alpha, float dx, float dy) {
    const uint32_t cb = __ballot_sync(0xffffffffu, contrib);
    if (cb == 0u) return;
    // a lane without a contributing pair replays alpha = G = 0: every partial is exactly 0 and
    // T, X, Bp keep their values
    alpha = contrib ? alpha : 0.f;
    G = contrib ? G : 0.f;
    float v[15];
    {
      const float4 g1 = s.r1[k];
      const float4 g2 = s.r2[k];
      const float4 g3 = s.r3[k];
      const float4 g4 = s.r4[k];
      const float om = 1.f - alpha;
      const float inv_om = recip_om(om);
      T = T * inv_om;                 // backward.cu:756
      const float w = alpha * T;
      const float wp = w * T;
      const float z = g4.w;
      const float2 kd = fma2(make_float2(g2.z, g2.w), make_float2(gc2, gd),
                             mul2(make_float2(g2.x, g2.y), make_float2(gc0, gc1)));
      const float x_tot = (kd.x + kd.y) + (z * (z * k2 - k1) + k0);
      const float2 pd = fma2(make_float2(g4.x, g4.y), make_float2(gp4, gp5),
                             fma2(make_float2(g3.z, g3.w), make_float2(gp2, gp3),
                                  mul2(make_float2(g3.x, g3.y), make_float2(gp0, gp1))));
      const float pi = (pd.x + pd.y) + g4.z * gp6;
      const float dL_dalpha = (x_tot - X) * T + (pi - 2.f * om * Bp) * (T * T) -
                              (T_final * inv_om) * bgdot;
      const float2 rec = fma2(make_float2(alpha, alpha), make_float2(x_tot, pi),
                              mul2(make_float2(om, om * om), make_float2(X, Bp)));
      X = rec.x;
      Bp = rec.y;
      const float h = g1.w * dL_dalpha * G;
      const float2 dxy = make_float2(dx, dy);
      const float2 hxy = mul2(make_float2(h, h), dxy);
      const float2 hxx = mul2(make_float2(hxy.x, hxy.x), dxy);
      v[0] = hxy.x; v[1] = hxy.y; v[2] = hxx.x; v[3] = hxx.y;
      v[4] = hxy.y * dy;
      v[5] = G * dL_dalpha;
      const float2 w2 = make_float2(w, w), wp2 = make_float2(wp, wp);
      const float2 v67 = mul2(w2, make_float2(gc0, gc1)), v89 = mul2(w2, make_float2(gc2, gd));
      const float2 vAB = mul2(wp2, make_float2(gA, gB)), vCS = mul2(wp2, make_float2(gp2, gS));
      v[6] = v67.x; v[7] = v67.y; v[8] = v89.x; v[9] = v89.y;
      v[10] = w * (2.f * z * k2 - k1);
      v[11] = vAB.x; v[12] = vAB.y; v[13] = vCS.x; v[14] = vCS.y;
    }
    // rows 0..14: the left half's values, rows 16..30 (+16 floats): the right half's
    float* col = red + half * (16 * REDH_STRIDE + 16) + (lane & 15u);
#pragma unroll
    for (int i = 0; i < 15; ++i) col[i * REDH_STRIDE] = v[i];
    __syncwarp();
    const float4* rp = reinterpret_cast<const float4*>(red + lane * REDH_STRIDE + half * 16u);
    const float4 q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3];
    const float2 s0 = add2(make_float2(q0.x, q0.y), make_float2(q0.z, q0.w));
    const float2 s1 = add2(make_float2(q1.x, q1.y), make_float2(q1.z, q1.w));
    const float2 s2 = add2(make_float2(q2.x, q2.y), make_float2(q2.z, q2.w));
    const float2 s3 = add2(make_float2(q3.x, q3.y), make_float2(q3.z, q3.w));
    const float2 st = add2(add2(s0, s1), add2(s2, s3));
    const float sum = st.x + st.y;
    __syncwarp();
    const uint32_t mine = half ? (cb >> 16) : (cb & 0xffffu);     // did this lane's half contribute at all
    if (mine != 0u && (lane & 15u) != 15u)
      atomicAdd(grad_rec + (size_t)s.id[k] * GFT_GRAD_FLOATS + (lane & 15u), sum);
  };
  auto walk_half = [&](const auto& s, int base, int m_w) {
    bool hitL = false, hitR = false;
    const int jj = (int)lane;
    if (jj < m_w) {
      const float4 g0 = s.r0[jj];
      const bool ymiss = g0.y + g0.w < patch_y0 || g0.y - g0.w > patch_y1;
      hitL = !(ymiss || g0.x + g0.z < patch_x0 || g0.x - g0.z > patch_x0 + 3.f);
      hitR = !(ymiss || g0.x + g0.z < patch_x0 + 4.f || g0.x - g0.z > patch_x1);
    }
    // positions >= the half's furthest contributor contribute nothing for that half
    const uint32_t hmL = __shfl_sync(0xffffffffu, hmax, 0), hmR = __shfl_sync(0xffffffffu, hmax, 16);
    hitL = hitL && (uint32_t)(base + jj) < hmL;
    hitR = hitR && (uint32_t)(base + jj) < hmR;
    uint32_t mL = __ballot_sync(0xffffffffu, hitL), mR = __ballot_sync(0xffffffffu, hitR);
    while (mL | mR) {
      const int kL1 = mL ? 31 - __clz(mL) : -1;
      if (mL) mL &= ~(1u << kL1);
      const int kR1 = mR ? 31 - __clz(mR) : -1;
      if (mR) mR &= ~(1u << kR1);
      const int kL2 = mL ? 31 - __clz(mL) : -1;
      if (mL) mL &= ~(1u << kL2);
      const int kR2 = mR ? 31 - __clz(mR) : -1;
      if (mR) mR &= ~(1u << kR2);
      const int k1 = half ? kR1 : kL1, k2 = half ? kR2 : kL2;
      float G1 = 0.f, G2 = 0.f, al1 = 0.f, al2 = 0.f, dx1 = 0.f, dy1 = 0.f, dx2 = 0.f, dy2 = 0.f;
      const bool c1 = k1 >= 0 && eval_pair(s, base, k1, G1, al1, dx1, dy1);
      const bool c2 = k2 >= 0 && eval_pair(s, base, k2, G2, al2, dx2, dy2);
      replay_half(s, max(k1, 0), c1, G1, al1, dx1, dy1);
      if (kL2 >= 0 || kR2 >= 0) replay_half(s, max(k2, 0), c2, G2, al2, dx2, dy2);
    }
  };

  if (RING) {
    const int nc = (n_eff + 31) >> 5;                 // chunks, processed from the last to the first
    auto request = [&](int q) {                       // whole warp; sequence number q <-> chunk nc-1-q
      const int slot = q % RING_SLOTS, fill = q / RING_SLOTS;
      if (fill > 0) mbar_wait(&s_empty[slot], (uint32_t)(fill - 1) & 1u);   // every warp left the old contents
      const int pos = (nc - 1 - q) * 32 + (int)lane;
      if (pos < n_eff) {
        const uint32_t* src = p.point_list + range.x + pos;
        const int g = (int)__ldg(src);
        const float4* r = recs + (size_t)g * (GFT_REC_FLOATS / 4);
        RingSlot& d = ring[slot];
        cp_async4_ca(&d.id[lane], src);
        cp_async16(&d.r0[lane], r + 0);
        cp_async16(&d.r1[lane], r + 1);
        cp_async16(&d.r2[lane], r + 2);
        cp_async16(&d.r3[lane], r + 3);
        cp_async16(&d.r4[lane], r + 4);
      }
      mbar_arrive_on_copies(&s_full[slot]);
    };
    if ((int)warp < nc) request((int)warp);           // chunks 0..7 of the sequence
    for (int q = 0; q < nc; ++q) {
      const int ahead = q + RING_LEAD;
      if (ahead < nc && (ahead & (WARPS - 1)) == (int)warp) request(ahead);
      const int slot = q % RING_SLOTS;
      mbar_wait(&s_full[slot], (uint32_t)(q / RING_SLOTS) & 1u);
      const int base = (nc - 1 - q) * 32;
      const int m_w = min(min(32, n_eff - base), (int)wmax - base);   // positions >= wmax: nothing for this warp
      if (m_w > 0) {
        if (HALF) walk_half(ring[slot], base, m_w);
        else walk(ring[slot], base, 0, m_w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[slot]);
    }
    return;
  }

  int cur = 0;
  if (nb > 0) stage(nb - 1, cur);
  for (int b = nb - 1; b >= 0; --b) {
    cp_async_wait_all();
    // batch b is in buf[cur]; everyone left buf[cur^1]
    __syncthreads();
    if (b > 0) stage(b - 1, cur ^ 1);      // gather the next (nearer) batch while this one is used
    const BwdBuf& s = buf[cur];
    const int base = b * BATCH;
    const int m = min(BATCH, n_eff - base);
    const int m_w = min(m, (int)wmax - base);  // positions >= wmax contribute nothing for this warp
    for (int c = ((m_w - 1) >> 5) << 5; c >= 0 && m_w > 0; c -= 32) walk(s, base, c, m_w);
    cur ^= 1;
  }
}

namespace {
template <int MINB, bool PRED, bool RING, bool HALF = false>
void launch_bwd_variant(const BlendBwdParams& p, cudaStream_t stream) {
  const int smem = 2 * (int)sizeof(BwdBufT<GFT_BLOCK>) + (GFT_BLOCK / 32) * (HALF ? REDH_FLOATS : RED_FLOATS) * 4;
  static unsigned long long smem_ok = 0;
  ensure_dynamic_smem(blend_bwd_kernel<MINB, PRED, RING, HALF>, smem, &smem_ok);
  blend_bwd_kernel<MINB, PRED, RING, HALF><<<p.T_total, GFT_BLOCK, smem, stream>>>(p);
}
}  // namespace

int blend_block_warps(int) { return 8; }

void launch_blend_bwd(const BlendBwdParams& p, cudaStream_t stream) {
  if (p.T_total <= 0) return;
  // Measured and rejected on B200 (c2 and c4 workloads): 64-register variants for one more resident
  // block per SM (+4 %), replays with few contributing pixels sent straight to the record with
  // per-lane reductions (+4 %), a warp-autonomous variant like the forward's (+6 %), half-tile
  // blocks (+-1 %), the shuffle butterfly instead of the shared-memory transpose (+7 %).
  // Options bwd_pred = 0 (branchy replay) and bwd_ring = 0 (block double buffer instead of the mbarrier
  // ring: +2.5 % at 640x480, +3.7 % at 1080p) for A/B runs.
  const bool pred = option(OPT_BWD_PRED) != 0, ring = option(OPT_BWD_RING) != 0;
  if (option(OPT_BLEND_HALF) != 0) {
    launch_bwd_variant<3, true, true, true>(p, stream);
  } else if (ring) {
    if (pred) launch_bwd_variant<3, true, true>(p, stream);
    else launch_bwd_variant<3, false, true>(p, stream);
  } else {
    if (pred) launch_bwd_variant<3, true, false>(p, stream);
    else launch_bwd_variant<3, false, false>(p, stream);
  }
  note_launches(1);
}

}  // namespace gft
