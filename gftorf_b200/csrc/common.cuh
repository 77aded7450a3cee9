// common.cuh — shared device helpers for the sm_100a RGB+ToF Gaussian rasterizer.
//
// Everything that decides an INTEGER result of the pipeline (radii, tile rectangles, the depth
// bits of the sort keys, n_contrib, pixels) is written with explicit-rounding intrinsics
// (__fmul_rn / __fadd_rn / __fmaf_rn / __fdiv_rn / __frcp_rn / __fsqrt_rn).  The operation order
// reproduces the dataflow nvcc 12.9 + ptxas generate for the reference expressions
// (cuda_rasterizer/forward.cu:128-206,281-342,524-543 and auxiliary.h:44-80 of
// submodules/diff-gaussian-rasterization-w-tof); intrinsics are never re-associated or
// contracted, so the result does not depend on how the surrounding code is compiled.
//
// Contraction rule that the reference binary follows (SASS-verified):
//   p1 + p2 + p3 (+ c)  ->  fma(p3, fma(p1, mul(p2))) (+ c as a separate add)
//   a*c - b*b           ->  fma(a, c, -mul(b,b))
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GFT_BLOCK 256
#define GFT_TILE_X 16
#define GFT_TILE_Y 16
#define GFT_REC_FLOATS 20   // blend record: 5 x float4 per Gaussian
#define GFT_GRAD_FLOATS 16  // blend-gradient record: 64 bytes per Gaussian (layout: blend_bwd.cu)

namespace gft {

// Spherical-harmonics constants, auxiliary.h:23-40 (constexpr scalars -> immediates in SASS)
constexpr float kSH_C0 = 0.28209479177387814f;
constexpr float kSH_C1 = 0.4886025119029199f;
constexpr float kSH_C2_0 = 1.0925484305920792f;
constexpr float kSH_C2_1 = -1.0925484305920792f;
constexpr float kSH_C2_2 = 0.31539156525252005f;
constexpr float kSH_C2_3 = -1.0925484305920792f;
constexpr float kSH_C2_4 = 0.5462742152960396f;
constexpr float kSH_C3_0 = -0.5900435899266435f;
constexpr float kSH_C3_1 = 2.890611442640554f;
constexpr float kSH_C3_2 = -0.4570457994644658f;
constexpr float kSH_C3_3 = 0.3731763325901154f;
constexpr float kSH_C3_4 = -0.4570457994644658f;
constexpr float kSH_C3_5 = 1.445305721320277f;
constexpr float kSH_C3_6 = -0.5900435899266435f;

#define GFT_PI_F 3.14159265358979323846f  // auxiliary.h:42

// a*b + c*d + e*f in the reference's contracted form: fma(e,f, fma(a,b, mul(c,d)))
__device__ __forceinline__ float dot3c(float a, float b, float c, float d, float e, float f) {
  return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}

// Column-major 4x4 applied to a point: m[0]x + m[4]y + m[8]z + m[12]  (auxiliary.h:61-80)
__device__ __forceinline__ float xform_row(const float* __restrict__ m, int r, float x, float y,
                                           float z) {
  return __fadd_rn(dot3c(x, m[r], y, m[r + 4], z, m[r + 8]), m[r + 12]);
}

struct Cov3 {
  float c0, c1, c2, c3, c4, c5;
};

// computeCov3D, forward.cu:172-206 (quaternion NOT normalised; (r,x,y,z) = (q0..q3)).
// For finite inputs the structural zeros of S drop out exactly: M[c][r] = fl(s_r * R[c][r]).
__device__ __forceinline__ Cov3 cov3d_from_scale_rot(float sx0, float sy0, float sz0, float mod,
                                                     float r, float x, float y, float z) {
  const float sx = __fmul_rn(mod, sx0), sy = __fmul_rn(mod, sy0), sz = __fmul_rn(mod, sz0);
  // Dataflow of the reference binary (SASS of preprocessCUDA, sm_100a): x*z, r*x, r*z, y*y, z*z are
  // rounded products; the other product of each sum/difference is fused into an FMA.
  const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
  const float xz_p_ry = __fmaf_rn(r, y, xz), xz_m_ry = __fmaf_rn(-r, y, xz);
  const float yz_m_rx = __fmaf_rn(y, z, -rx), yz_p_rx = __fmaf_rn(y, z, rx);
  const float xy_m_rz = __fmaf_rn(x, y, -rz), xy_p_rz = __fmaf_rn(x, y, rz);
  const float xx_zz = __fmaf_rn(x, x, zz);
  const float xx_yy = __fmaf_rn(x, x, yy);
  const float yy_zz = __fadd_rn(yy, zz);
  // glm::mat3 R (column-major fill): R[0]=(R00,R01,R02) ...
  const float R00 = __fsub_rn(1.f, __fadd_rn(yy_zz, yy_zz));
  const float R01 = __fadd_rn(xy_m_rz, xy_m_rz);
  const float R02 = __fadd_rn(xz_p_ry, xz_p_ry);
  const float R10 = __fadd_rn(xy_p_rz, xy_p_rz);
  const float R11 = __fsub_rn(1.f, __fadd_rn(xx_zz, xx_zz));
  const float R12 = __fadd_rn(yz_m_rx, yz_m_rx);
  const float R20 = __fadd_rn(xz_m_ry, xz_m_ry);
  const float R21 = __fadd_rn(yz_p_rx, yz_p_rx);
  const float R22 = __fsub_rn(1.f, __fadd_rn(xx_yy, xx_yy));
  // M = S * R
  const float M00 = __fmul_rn(sx, R00), M01 = __fmul_rn(sy, R01), M02 = __fmul_rn(sz, R02);
  const float M10 = __fmul_rn(sx, R10), M11 = __fmul_rn(sy, R11), M12 = __fmul_rn(sz, R12);
  const float M20 = __fmul_rn(sx, R20), M21 = __fmul_rn(sy, R21), M22 = __fmul_rn(sz, R22);
  // Sigma = transpose(M) * M, glm product order (type_mat3x3.inl:510-518)
  Cov3 c;
  c.c0 = dot3c(M00, M00, M01, M01, M02, M02);
  c.c1 = dot3c(M10, M00, M11, M01, M12, M02);
  c.c2 = dot3c(M20, M00, M21, M01, M22, M02);
  c.c3 = dot3c(M10, M10, M11, M11, M12, M12);
  c.c4 = dot3c(M20, M10, M21, M11, M22, M12);
  c.c5 = dot3c(M20, M20, M21, M21, M22, M22);
  return c;
}

// The 2x3 non-zero part of T = W * J used by computeCov2D (forward.cu:128-167,
// backward.cu:287-313).  T0* = T[0][*], T1* = T[1][*] in glm indexing.
struct Tmat {
  float T00, T01, T02, T10, T11, T12;
};

__device__ __forceinline__ Tmat ewa_T(const float* __restrict__ V, float tx, float ty, float tz,
                                      float focal_x, float focal_y, float tan_fovx,
                                      float tan_fovy, float* cx_out = nullptr,
                                      float* cy_out = nullptr, float* txtz_out = nullptr,
                                      float* tytz_out = nullptr) {
  const float limx = __fmul_rn(tan_fovx, 1.3f);
  const float limy = __fmul_rn(tan_fovy, 1.3f);
  const float txtz = __fdiv_rn(tx, tz);
  const float tytz = __fdiv_rn(ty, tz);
  const float cx = fminf(limx, fmaxf(-limx, txtz));
  const float cy = fminf(limy, fmaxf(-limy, tytz));
  if (cx_out) *cx_out = cx;
  if (cy_out) *cy_out = cy;
  if (txtz_out) *txtz_out = txtz;
  if (tytz_out) *tytz_out = tytz;
  const float tz2 = __fmul_rn(tz, tz);
  const float J00 = __fdiv_rn(focal_x, tz);
  const float J02 = __fdiv_rn(__fmul_rn(focal_x, __fmul_rn(cx, -tz)), tz2);
  const float J11 = __fdiv_rn(focal_y, tz);
  const float J12 = __fdiv_rn(__fmul_rn(focal_y, __fmul_rn(cy, -tz)), tz2);
  Tmat t;
  t.T00 = __fmaf_rn(V[2], J02, __fmul_rn(V[0], J00));
  t.T01 = __fmaf_rn(V[6], J02, __fmul_rn(V[4], J00));
  t.T02 = __fmaf_rn(V[10], J02, __fmul_rn(V[8], J00));
  t.T10 = __fmaf_rn(V[2], J12, __fmul_rn(J11, V[1]));
  t.T11 = __fmaf_rn(V[6], J12, __fmul_rn(J11, V[5]));
  t.T12 = __fmaf_rn(V[10], J12, __fmul_rn(J11, V[9]));
  return t;
}

// cov = transpose(T) * transpose(Vrk) * T, returns (cov[0][0]+0.3, cov[0][1], cov[1][1]+0.3)
__device__ __forceinline__ float3 ewa_cov2d(const Tmat& t, const Cov3& v) {
  const float B00 = dot3c(t.T00, v.c0, t.T01, v.c1, t.T02, v.c2);
  const float B01 = dot3c(t.T10, v.c0, t.T11, v.c1, t.T12, v.c2);
  const float B10 = dot3c(t.T00, v.c1, t.T01, v.c3, t.T02, v.c4);
  const float B11 = dot3c(t.T10, v.c1, t.T11, v.c3, t.T12, v.c4);
  const float B20 = dot3c(t.T00, v.c2, t.T01, v.c4, t.T02, v.c5);
  const float B21 = dot3c(t.T10, v.c2, t.T11, v.c4, t.T12, v.c5);
  float3 c;
  c.x = __fadd_rn(dot3c(t.T00, B00, t.T01, B10, t.T02, B20), 0.3f);
  c.y = dot3c(t.T00, B01, t.T01, B11, t.T02, B21);
  c.z = __fadd_rn(dot3c(t.T10, B01, t.T11, B11, t.T12, B21), 0.3f);
  return c;
}

// ndc2Pix, auxiliary.h:44-47 — evaluated in double with a double FMA, rounded once to float.
__device__ __forceinline__ float ndc2pix(float v, int S) {
  return __double2float_rn(
      __dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5));
}

// getRect, auxiliary.h:49-59.  `radius` is the int-converted radius; grid in tiles.
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy,
                                          uint32_t& x0, uint32_t& y0, uint32_t& x1, uint32_t& y1) {
  const float rf = (float)radius;
  x0 = min((uint32_t)gx, (uint32_t)max(0, (int)__fmul_rn(__fsub_rn(px, rf), 0.0625f)));
  y0 = min((uint32_t)gy, (uint32_t)max(0, (int)__fmul_rn(__fsub_rn(py, rf), 0.0625f)));
  x1 = min((uint32_t)gx,
           (uint32_t)max(0, (int)__fmul_rn(
                                __fadd_rn(__fadd_rn(__fadd_rn(px, rf), 16.f), -1.f), 0.0625f)));
  y1 = min((uint32_t)gy,
           (uint32_t)max(0, (int)__fmul_rn(
                                __fadd_rn(__fadd_rn(__fadd_rn(py, rf), 16.f), -1.f), 0.0625f)));
}

// Gaussian falloff exponent of one (pixel, Gaussian) pair, forward.cu:524-527 as compiled:
//   u = dy*(dy*C); v = dy*(dx*B); q = fma(dx, dx*A, u); power = fma(q, -0.5, -v)
__device__ __forceinline__ float pair_power(float dx, float dy, float A, float B, float C) {
  const float u = __fmul_rn(dy, __fmul_rn(dy, C));
  const float v = __fmul_rn(dy, __fmul_rn(dx, B));
  const float q = __fmaf_rn(dx, __fmul_rn(dx, A), u);
  return __fmaf_rn(q, -0.5f, -v);
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2) ------------------------------
// One instruction, two IEEE round-to-nearest results — each component is exactly what the scalar
// instruction gives, so bit-pinned arithmetic may use them.  The blend kernels are bound by
// instruction issue, and their accumulators come in natural pairs (two halves of a 128-bit load).
// (Measured: packing the accumulations and the reduction tree pays, -5 % / -5 % on the blend
// kernels; packing the short dependent chain of pair_power() does not, +3 % on the forward.)
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

// ---- warp helpers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Visits every tile of a Gaussian's tile rectangle [rx0,rx1) x [ry0, ry0 + tiles/(rx1-rx0)):
// rectangles of up to 32 tiles by their own lane, larger ones by the whole warp (the reference's
// initial splats have a median radius of ~50 px and a few cover thousands of tiles; one thread
// looping over all of them stalls its warp, rasterizer_impl.cu:90-110).  Must be called by all 32
// lanes; lanes with nothing to emit pass tiles = 0.  `payload` travels with the rectangle.
template <typename F>
__device__ __forceinline__ void for_each_tile(uint32_t rx0, uint32_t ry0, uint32_t rx1, uint32_t tiles,
                                              uint32_t payload, uint32_t lane, F&& f) {
  const uint32_t w = rx1 - rx0;
  const uint32_t big = __ballot_sync(0xffffffffu, tiles > 32u);
  if (tiles != 0u && tiles <= 32u) {
    uint32_t tx = rx0, ty = ry0;
    for (uint32_t k = 0; k < tiles; ++k) {
      f(tx, ty, payload);
      if (++tx == rx1) { tx = rx0; ++ty; }
    }
  }
  uint32_t m = big;
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t bx0 = __shfl_sync(0xffffffffu, rx0, src), by0 = __shfl_sync(0xffffffffu, ry0, src);
    const uint32_t bw = __shfl_sync(0xffffffffu, w, src), bn = __shfl_sync(0xffffffffu, tiles, src);
    const uint32_t bp = __shfl_sync(0xffffffffu, payload, src);
    for (uint32_t k = lane; k < bn; k += 32u) f(bx0 + k % bw, by0 + k / bw, bp);
  }
}

// The same walk in two phases, four tiles at a time: `acquire(tx, ty, payload)` returns two 32-bit
// words (a loaded base and the old value of an atomic cursor) that `commit(payload, payload2, words)` uses.
// The four acquires are issued before the first word is consumed, so four atomics are in flight
// per lane instead of one (the scatter is bound by the round trip of the returning atomic).
template <typename A, typename C>
__device__ __forceinline__ void for_each_tile_2phase(uint32_t rx0, uint32_t ry0, uint32_t rx1, uint32_t tiles,
                                                     uint32_t payload, uint32_t payload2, uint32_t lane, A&& acquire,
                                                     C&& commit) {
  const uint32_t w = rx1 - rx0;
  const uint32_t big = __ballot_sync(0xffffffffu, tiles > 32u);
  if (tiles != 0u && tiles <= 32u) {
    uint32_t tx = rx0, ty = ry0;
    for (uint32_t k = 0; k < tiles; k += 4u) {
      uint2 h[4];
#pragma unroll
      for (uint32_t u = 0; u < 4u; ++u) {
        h[u] = (k + u < tiles) ? acquire(tx, ty, payload) : make_uint2(0u, 0u);
        if (++tx == rx1) { tx = rx0; ++ty; }
      }
#pragma unroll
      for (uint32_t u = 0; u < 4u; ++u)
        if (k + u < tiles) commit(payload, payload2, h[u]);
    }
  }
  uint32_t m = big;
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t bx0 = __shfl_sync(0xffffffffu, rx0, src), by0 = __shfl_sync(0xffffffffu, ry0, src);
    const uint32_t bw = __shfl_sync(0xffffffffu, w, src), bn = __shfl_sync(0xffffffffu, tiles, src);
    const uint32_t bp = __shfl_sync(0xffffffffu, payload, src), bp2 = __shfl_sync(0xffffffffu, payload2, src);
    for (uint32_t k = lane; k < bn; k += 128u) {
      uint2 h[4];
#pragma unroll
      for (uint32_t u = 0; u < 4u; ++u) {
        const uint32_t kk = k + 32u * u;
        h[u] = (kk < bn) ? acquire(bx0 + kk % bw, by0 + kk / bw, bp) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (uint32_t u = 0; u < 4u; ++u)
        if (k + 32u * u < bn) commit(bp, bp2, h[u]);
    }
  }
}

// Lanes of the warp whose 8-bit digit equals this lane's, among the lanes with `ok` set: eight
// ballots, one per digit bit (independent, pipelined).  The MATCH.ANY instruction gives the same
// mask; measured on B200 the ballots make the per-tile radix sort 12 % faster (0.104 vs 0.117 ms
// at 2.4 M instances) — option sort_match = 1 selects MATCH.ANY for A/B runs.
__device__ __forceinline__ uint32_t match_digit8(uint32_t d, bool ok) {
  uint32_t m = __ballot_sync(0xffffffffu, ok);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

// ---- warp-cooperative staging of per-Gaussian rows ------------------------------------------
// The SH tensors are [P][ROW] row-major, so the rows of a warp's 32 consecutive Gaussians are one
// contiguous chunk of global memory.  These helpers move the chunk between global memory (16-byte
// accesses over whole sectors, 6 / 4 of them in flight per lane: measured -10 ... -13 % on the
// preprocess kernels against 4-byte accesses "lane l touches element l, l+32, ...") and a
// warp-private shared buffer with row stride ROW+1 floats, in which lane l then walks row l
// without bank conflicts ((ROW+1) is odd).
// Which 16-byte piece of the chunk a lane moves in iteration `it`: row r (0..31) and first column c
// (a multiple of 4).  Chosen so that (1) a warp instruction touches whole 32-byte sectors of global
// memory (4 rows x 128 contiguous bytes, or 8 rows x 64) and (2) the four scalar shared-memory
// accesses that follow are conflict-free with the row stride ROW + 1 (bank = (17 r + c) mod 32 for
// ROW = 48, (r + c) mod 32 for ROW = 32).  Twelve (ROW = 48) or eight (ROW = 32) iterations cover
// the 32 rows.
template <int ROW>
__device__ __forceinline__ void stage_piece(int it, uint32_t lane, int& r, int& c) {
  static_assert(ROW == 48 || ROW == 32, "row widths of the SH tensors: 16 x 3 and 16 x 2");
  if (ROW == 32 || it < 8) {
    r = 4 * it + (int)(lane & 3u);
    c = 4 * (int)(lane >> 2);
  } else {
    r = (int)(lane & 16u) + 4 * (it - 8) + (int)(lane & 3u);
    c = 32 + 4 * (int)((lane >> 2) & 3u);
  }
}

template <int ROW>
__device__ __forceinline__ void warp_stage_in(const float* __restrict__ gbase, int nrows,
                                              float* sbuf, uint32_t lane) {
  constexpr int ITERS = ROW / 4;          // 16-byte loads per lane
  constexpr int CH = ROW == 48 ? 6 : 4;   // loads in flight per lane (3 / 2 KB per warp)
#pragma unroll
  for (int h = 0; h < ITERS; h += CH) {
    float4 v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      int r, c;
      stage_piece<ROW>(h + k, lane, r, c);
      if (r < nrows) v[k] = __ldg(reinterpret_cast<const float4*>(gbase + r * ROW + c));
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      int r, c;
      stage_piece<ROW>(h + k, lane, r, c);
      if (r < nrows) {
        float* d = sbuf + r * (ROW + 1) + c;
        d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
      }
    }
  }
}
// Gradient store modes of the backward: 0 overwrite, 1 add (plain read-modify-write: one view at
// a time per buffer), 2 atomic add (several views may run concurrently into one bucket).
template <int ACC>
__device__ __forceinline__ void acc_store(float* p, float v) {
  if (ACC == 0) *p = v;
  else if (ACC == 1) *p += v;
  else atomicAdd(p, v);
}

template <int ROW, int ACC = 0>
__device__ __forceinline__ void warp_stage_out(float* __restrict__ gbase, int nrows,
                                               const float* sbuf, uint32_t lane) {
  constexpr int ITERS = ROW / 4;
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    int r, c;
    stage_piece<ROW>(it, lane, r, c);
    if (r < nrows) {
      const float* sp = sbuf + r * (ROW + 1) + c;
      float* g = gbase + r * ROW + c;
      if (ACC == 0) {
        *reinterpret_cast<float4*>(g) = make_float4(sp[0], sp[1], sp[2], sp[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc_store<ACC>(g + j, sp[j]);
      }
    }
  }
}
#define GFT_STAGE_FLOATS_PER_WARP (32 * 49)

// 64-bit status words of the decoupled look-back scan.  flag and value travel in ONE word, so
// relaxed gpu-scope accesses suffice (an acquire poll costs an L1 invalidate every time).
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace gft
