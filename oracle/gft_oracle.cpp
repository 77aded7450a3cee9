// gft_oracle.cpp — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's RGB+ToF Gaussian
// rasterizer (forward + backward) and of distCUDA2, in float32, as plain loops.
//
// It is the checker of the GPU-less tests and the "port" CPU baseline of bench.py; the product
// path (gftorf_b200/) never links, loads or calls it.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / reference legs may.
//
// PARITY PINNING.  The reference ships no tests, golden vectors or CPU implementation for this
// path, so this restatement cannot be pinned against fixtures of the
// reference's own.  It is pinned instead against OUTPUTS OF THE REFERENCE ITSELF: the unmodified
// reference kernels compiled into oracle/_ref/libgftorf_ref.so (oracle/Makefile, oracle/ref_shim.cu)
// were run on a B200 on the seeded inputs of tests/golden/make_golden.py and their outputs are
// committed under tests/golden/; tests/test_oracle_golden.py holds this file to them (integers
// bit-exact except where CUDA's expf differs from libm's in the last ulp, see below).
//
// What each part follows (paths relative to submodules/diff-gaussian-rasterization-w-tof):
//   preprocess_one()   cuda_rasterizer/forward.cu:251-419 with helpers :20-71 (SH->RGB), :73-125
//                      (SH->phase,amp), :128-167 (EWA cov2D), :172-206 (cov3D),
//                      auxiliary.h:44-59 (ndc2Pix, getRect), :61-80 (transforms), :152-179 (frustum)
//   binning            rasterizer_impl.cu:72-113 (keys), :331-339 (stable sort on tile|depth),
//                      :118-140 (tile ranges), :35-50 (tile bit count)
//   blend_tile_fwd()   forward.cu:447-675
//   blend_tile_bwd()   backward.cu:632-888
//   cov2d_bwd()        backward.cu:276-394
//   preprocess_bwd()   backward.cu:491-605 with :20-139, :143-260 (SH backward), :399-462 (cov3D)
//   knn                submodules/simple-knn/simple_knn.cu:119-183 (exact 3-NN; brute force here,
//                      the result is independent of the visiting order)
//
// Arithmetic.  Compiled with -ffp-contract=off; wherever a rounding decides an INTEGER result
// (radii, tile rectangles, depth key bits, the alpha chain behind n_contrib and pixels) the
// fused multiply-adds the reference BINARY performs are written out with fmaf(), in the dataflow
// read from its SASS.  libm's expf/sinf/cosf are not CUDA's: they may differ
// in the last ulp, which can flip an alpha threshold for a rare pixel — the tests bound the count.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/gftorf.h"

namespace {

constexpr float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
constexpr float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                         -1.0925484305920792f, 0.5462742152960396f};
constexpr float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                         0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                         -0.5900435899266435f};
constexpr float PI_F = 3.14159265358979323846f;
constexpr int TILE = 16, BATCH = 256;

// a*b + c*d + e*f as the reference binary evaluates it: second product rounded alone, first and
// third fused
inline float dot3(float a, float b, float c, float d, float e, float f) {
  return fmaf(e, f, fmaf(a, b, c * d));
}
inline float xf(const float* m, int r, float x, float y, float z) {  // auxiliary.h:61-80
  return dot3(x, m[r], y, m[r + 4], z, m[r + 8]) + m[r + 12];
}

struct State {  // what the reference keeps in its geometry / binning / image buffers
  int P = 0, W = 0, H = 0, gx = 0, gy = 0, R = 0;
  std::vector<float> depths, ndc, dists, means2D, cov3D, conic_opacity, rgb, ria, pa;
  std::vector<uint8_t> clamped, clamped_p;
  std::vector<uint32_t> tiles_touched, point_offsets, rect;  // rect: x0 y0 x1 y1
  std::vector<uint64_t> keys;
  std::vector<uint32_t> point_list;
  std::vector<uint32_t> ranges;  // [T][2]
  std::vector<float> final_T, wz, wz2;
  std::vector<uint32_t> n_contrib;
};

thread_local std::string g_err;

// SH basis values multiplying coefficients 1..15 for unit direction (x,y,z), forward.cu:37-59 in
// the reference binary's dataflow (SASS): products rounded, except the fused polynomial terms
// 3xx-yy, 2zz-3xx-3yy and xx-3yy.
void sh_basis(int deg, float x, float y, float z, float* b) {
  if (deg < 1) return;
  b[1] = -(C1 * y); b[2] = C1 * z; b[3] = -(C1 * x);
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  const float zz2 = zz + zz, xx_yy = xx - yy;
  b[4] = xy * C2[0]; b[5] = yz * C2[1]; b[6] = ((zz2 - xx) - yy) * C2[2];
  b[7] = xz * C2[3]; b[8] = xx_yy * C2[4];
  if (deg < 3) return;
  const float t11 = fmaf(zz, 4.f, -xx) - yy;
  b[9] = (y * C3[0]) * fmaf(xx, 3.f, -yy); b[10] = (xy * C3[1]) * z;
  b[11] = (y * C3[2]) * t11;
  b[12] = (z * C3[3]) * fmaf(yy, -3.f, fmaf(xx, -3.f, zz2));
  b[13] = t11 * (x * C3[4]); b[14] = xx_yy * (z * C3[5]);
  b[15] = (x * C3[6]) * fmaf(yy, -3.f, xx);
}

struct Tm { float T00, T01, T02, T10, T11, T12; };

// upper 2x3 of T = W*J, forward.cu:134-153 / backward.cu:287-313, pinned dataflow (A.8)
Tm ewa_T(const float* V, float tx, float ty, float tz, float fx, float fy, float tanx, float tany) {
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  const float cx = std::fmin(limx, std::fmax(-limx, tx / tz));
  const float cy = std::fmin(limy, std::fmax(-limy, ty / tz));
  const float tz2 = tz * tz;
  const float J00 = fx / tz, J02 = (fx * (cx * -tz)) / tz2;
  const float J11 = fy / tz, J12 = (fy * (cy * -tz)) / tz2;
  Tm t;
  t.T00 = fmaf(V[2], J02, V[0] * J00); t.T01 = fmaf(V[6], J02, V[4] * J00);
  t.T02 = fmaf(V[10], J02, V[8] * J00);
  t.T10 = fmaf(V[2], J12, J11 * V[1]); t.T11 = fmaf(V[6], J12, J11 * V[5]);
  t.T12 = fmaf(V[10], J12, J11 * V[9]);
  return t;
}

void cov2d(const Tm& t, const float* v, float& a, float& b, float& c) {  // forward.cu:155-166
  const float B00 = dot3(t.T00, v[0], t.T01, v[1], t.T02, v[2]);
  const float B01 = dot3(t.T10, v[0], t.T11, v[1], t.T12, v[2]);
  const float B10 = dot3(t.T00, v[1], t.T01, v[3], t.T02, v[4]);
  const float B11 = dot3(t.T10, v[1], t.T11, v[3], t.T12, v[4]);
  const float B20 = dot3(t.T00, v[2], t.T01, v[4], t.T02, v[5]);
  const float B21 = dot3(t.T10, v[2], t.T11, v[4], t.T12, v[5]);
  a = dot3(t.T00, B00, t.T01, B10, t.T02, B20) + 0.3f;
  b = dot3(t.T00, B01, t.T01, B11, t.T02, B21);
  c = dot3(t.T10, B01, t.T11, B11, t.T12, B21) + 0.3f;
}

void cov3d(const float* s0, float mod, const float* q, float* out) {  // forward.cu:172-206
  const float sx = mod * s0[0], sy = mod * s0[1], sz = mod * s0[2];
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  // reference binary (SASS): x*z, r*x, r*z, y*y, z*z rounded; the partner product is fused
  const float yy = y * y, zz = z * z, xz = x * z, rx = r * x, rz = r * z;
  const float xz_p_ry = fmaf(r, y, xz), xz_m_ry = fmaf(-r, y, xz);
  const float yz_m_rx = fmaf(y, z, -rx), yz_p_rx = fmaf(y, z, rx);
  const float xy_m_rz = fmaf(x, y, -rz), xy_p_rz = fmaf(x, y, rz);
  const float xx_zz = fmaf(x, x, zz), xx_yy = fmaf(x, x, yy), yy_zz = yy + zz;
  const float R00 = 1.f - (yy_zz + yy_zz), R01 = xy_m_rz + xy_m_rz, R02 = xz_p_ry + xz_p_ry;
  const float R10 = xy_p_rz + xy_p_rz, R11 = 1.f - (xx_zz + xx_zz), R12 = yz_m_rx + yz_m_rx;
  const float R20 = xz_m_ry + xz_m_ry, R21 = yz_p_rx + yz_p_rx, R22 = 1.f - (xx_yy + xx_yy);
  const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
  const float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
  const float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
  out[0] = dot3(M00, M00, M01, M01, M02, M02);
  out[1] = dot3(M10, M00, M11, M01, M12, M02);
  out[2] = dot3(M20, M00, M21, M01, M22, M02);
  out[3] = dot3(M10, M10, M11, M11, M12, M12);
  out[4] = dot3(M20, M10, M21, M11, M22, M12);
  out[5] = dot3(M20, M20, M21, M21, M22, M22);
}

inline float ndc2pix(float v, int S) {  // auxiliary.h:44-47, double with one fused step
  return (float)(std::fma((double)v + 1.0, (double)S, -1.0) * 0.5);
}

inline float pair_power(float dx, float dy, float A, float B, float C) {  // A.9
  const float u = dy * (dy * C), v = dy * (dx * B);
  const float q = fmaf(dx, dx * A, u);
  return fmaf(q, -0.5f, -v);
}

int tile_bits(uint32_t n) {  // rasterizer_impl.cu:35-50: bit length, 1 for 0
  int b = 0;
  while (b < 32 && (n >> b)) ++b;
  return b ? b : 1;
}

void preprocess_one(const GftForwardArgs& a, State& s, int i, float fx, float fy, float d2p) {
  a.radii[i] = 0;
  s.tiles_touched[i] = 0;
  a.pixels[i] = 0.f;
  const float* V = a.viewmatrix;
  const float* PM = a.projmatrix;
  const float px = a.means3D[3 * i], py = a.means3D[3 * i + 1], pz = a.means3D[3 * i + 2];
  const float vz = xf(V, 2, px, py, pz);
  if (vz < a.near_n || vz > a.far_n) return;  // auxiliary.h:167-177, z only
  const float vx = xf(V, 0, px, py, pz), vy = xf(V, 1, px, py, pz);
  const float hx = xf(PM, 0, px, py, pz), hy = xf(PM, 1, px, py, pz), hw = xf(PM, 3, px, py, pz);
  const float p_w = 1.0f / (hw + 0.0000001f);
  const float projx = hx * p_w, projy = hy * p_w;

  const float* c3;
  if (a.cov3D_precomp) {
    c3 = a.cov3D_precomp + 6 * (size_t)i;
  } else {
    cov3d(a.scales + 3 * (size_t)i, a.scale_modifier, a.rotations + 4 * (size_t)i, &s.cov3D[6 * (size_t)i]);
    c3 = &s.cov3D[6 * (size_t)i];
  }
  const Tm T = ewa_T(V, vx, vy, vz, fx, fy, a.tan_fovx, a.tan_fovy);
  float ca, cb, cc;
  cov2d(T, c3, ca, cb, cc);
  const float det = fmaf(ca, cc, -(cb * cb));
  if (det == 0.0f) return;
  const float det_inv = 1.f / det;
  const float conA = cc * det_inv, conB = cb * -det_inv, conC = ca * det_inv;
  const float mid = (ca + cc) * 0.5f;
  const float sq = std::sqrt(std::fmax(fmaf(mid, mid, -det), 0.1f));
  const float lam = std::fmax(mid + sq, mid - sq);
  const int radius = (int)std::ceil(std::sqrt(lam) * 3.f);
  const float pixx = ndc2pix(projx, a.width), pixy = ndc2pix(projy, a.height);
  const float rf = (float)radius;  // getRect, auxiliary.h:49-59
  auto clampi = [](int v, int hi) { return (uint32_t)std::min<int64_t>((uint32_t)hi, (uint32_t)std::max(0, v)); };
  const uint32_t x0 = clampi((int)((pixx - rf) * 0.0625f), s.gx);
  const uint32_t y0 = clampi((int)((pixy - rf) * 0.0625f), s.gy);
  const uint32_t x1 = clampi((int)((((pixx + rf) + 16.f) - 1.f) * 0.0625f), s.gx);
  const uint32_t y1 = clampi((int)((((pixy + rf) + 16.f) - 1.f) * 0.0625f), s.gy);
  if ((x1 - x0) * (y1 - y0) == 0) return;

  float basis[16];
  float dx = 0, dy = 0, dz = 0;
  if (a.shs || a.shs_p) {
    dx = px - a.campos[0]; dy = py - a.campos[1]; dz = pz - a.campos[2];
    const float len = std::sqrt(dot3(dx, dx, dy, dy, dz, dz));
    dx /= len; dy /= len; dz /= len;
    sh_basis(a.sh_degree, dx, dy, dz, basis);
  }
  const int nco = (a.sh_degree + 1) * (a.sh_degree + 1);
  float* rgb = &s.rgb[3 * (size_t)i];
  if (a.colors_precomp) for (int c = 0; c < 3; ++c) rgb[c] = a.colors_precomp[3 * (size_t)i + c];
  if (a.shs) {
    const float* sh = a.shs + (size_t)i * a.M * 3;
    for (int c = 0; c < 3; ++c) {
      float r = C0 * sh[c];
      for (int k = 1; k < nco; ++k) r = fmaf(basis[k], sh[3 * k + c], r);
      r += 0.5f;
      s.clamped[3 * (size_t)i + c] = r < 0.f;
      rgb[c] = std::fmax(r, 0.f);
    }
  }
  const float dist = std::sqrt(dot3(vx, vx, vy, vy, vz, vz));  // SASS: y^2 rounded, x^2, z^2 fused
  const float ndc = (1.f - a.near_n / dist) * (a.far_n / (a.far_n - a.near_n));
  const float factor = 1.0f / (dist * dist);
  float* ria = &s.ria[7 * (size_t)i];
  for (int c = 0; c < 7; ++c) ria[c] = 0.f;  // uninitialised in the reference when no phasor input
  bool have = false;
  float phase = 0, amp = 0;
  if (a.phasors_precomp) {  // forward.cu:365-387 (no phase_offset on this path)
    phase = dist * d2p;
    const float p0 = a.phasors_precomp[2 * (size_t)i], p1 = a.phasors_precomp[2 * (size_t)i + 1];
    s.pa[2 * (size_t)i] = p0; s.pa[2 * (size_t)i + 1] = p1;
    if (a.use_view_dependent_phase) phase += p0;
    amp = p1; have = true;
  }
  if (a.shs_p) {  // forward.cu:73-125, 389-407
    const float* sp = a.shs_p + (size_t)i * a.M_p * 2;
    float q0 = C0 * sp[0], q1 = C0 * sp[1];
    for (int k = 1; k < nco; ++k) { q0 = fmaf(basis[k], sp[2 * k], q0); q1 = fmaf(basis[k], sp[2 * k + 1], q1); }
    q1 += 0.5f;
    q0 = ((q0 + 0.5f) - 0.5f) - C0 * sp[0];
    s.clamped_p[i] = q1 < 0.f;
    if (q1 < 0.f) q1 = 0.f;
    s.pa[2 * (size_t)i] = q0; s.pa[2 * (size_t)i + 1] = q1;
    phase = fmaf(dist, d2p, a.phase_offset);
    if (a.use_view_dependent_phase) phase += q0;
    amp = q1; have = true;
  }
  if (have) {
    const float cs = std::cos(phase), sn = std::sin(phase), dc = a.dc_offset;
    ria[0] = cs * amp * factor; ria[1] = sn * amp * factor; ria[2] = amp * factor;
    ria[3] = (cs + dc) * amp * factor; ria[4] = (-cs + dc) * amp * factor;
    ria[5] = (sn + dc) * amp * factor; ria[6] = (-sn + dc) * amp * factor;
  }
  s.dists[i] = dist; s.depths[i] = vz; s.ndc[i] = ndc;
  a.radii[i] = radius;
  s.means2D[2 * (size_t)i] = pixx; s.means2D[2 * (size_t)i + 1] = pixy;
  float* co = &s.conic_opacity[4 * (size_t)i];
  co[0] = conA; co[1] = conB; co[2] = conC; co[3] = a.opacities[i];
  uint32_t* rc = &s.rect[4 * (size_t)i];
  rc[0] = x0; rc[1] = y0; rc[2] = x1; rc[3] = y1;
  s.tiles_touched[i] = (y1 - y0) * (x1 - x0);
}

inline float bgv(const float* bg, int mode, int ch, size_t HW, size_t pix) {
  return mode == 0 ? bg[ch * HW + pix] : bg[ch];
}

void blend_tile_fwd(const GftForwardArgs& a, State& s, int tile, std::vector<float>& pix_counts) {
  const int tx = tile % s.gx, ty = tile / s.gx;
  const uint32_t r0 = s.ranges[2 * tile], r1 = s.ranges[2 * tile + 1];
  const int n = (int)(r1 - r0);
  const size_t HW = (size_t)a.width * a.height;
  bool done[TILE * TILE];
  int ndone = 0;
  struct Px { float T, C[3], Pq[7], D, A, DD, DD_D, DD_D2, WD[3]; uint32_t contributor, last; int gs; };
  static thread_local std::vector<Px> px(TILE * TILE);
  for (int t = 0; t < TILE * TILE; ++t) {
    const int x = tx * TILE + t % TILE, y = ty * TILE + t / TILE;
    done[t] = !(x < a.width && y < a.height);
    ndone += done[t];
    Px& p = px[t];
    std::memset(&p, 0, sizeof(Px));
    p.T = 1.0f;
  }
  // batches of 256: a pixel that is done stops; the tile stops at a batch boundary once all are
  // done (forward.cu:497-502) — per pixel this is the same as stopping at `done`.
  for (int base = 0; base < n && ndone < TILE * TILE; base += BATCH) {
    const int m = std::min(BATCH, n - base);
    for (int t = 0; t < TILE * TILE; ++t) {
      if (done[t]) continue;
      Px& p = px[t];
      const float pfx = (float)(tx * TILE + t % TILE), pfy = (float)(ty * TILE + t / TILE);
      for (int j = 0; j < m; ++j) {
        p.contributor++;
        const uint32_t g = s.point_list[r0 + base + j];
        const float* co = &s.conic_opacity[4 * (size_t)g];
        const float dx = s.means2D[2 * (size_t)g] - pfx, dy = s.means2D[2 * (size_t)g + 1] - pfy;
        const float power = pair_power(dx, dy, co[0], co[1], co[2]);
        if (power > 0.0f) continue;
        const float alpha = std::fmin(0.99f, co[3] * std::exp(power));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = p.T * (1 - alpha);
        if (test_T < 0.0001f) { done[t] = true; ++ndone; break; }
        const float w = alpha * p.T, wp = w * p.T;
        const float* rgb = &s.rgb[3 * (size_t)g];
        const float* ria = &s.ria[7 * (size_t)g];
        for (int c = 0; c < 3; ++c) p.C[c] += rgb[c] * w;
        for (int c = 0; c < 7; ++c) p.Pq[c] += ria[c] * wp;
        p.D += s.dists[g] * w;
        if (p.gs < 1) { p.WD[0] = alpha; p.WD[1] = s.dists[g]; p.WD[2] = ria[2]; }
        p.gs += 1;
        const float z = s.ndc[g];
        p.DD += w * (z * z * p.A - 2.0f * z * p.DD_D + p.DD_D2);
        p.DD_D += w * z;
        p.DD_D2 += w * z * z;
        p.A += w;
        p.T = test_T;
        p.last = p.contributor;
        pix_counts[g] += 1.0f;
      }
    }
  }
  for (int t = 0; t < TILE * TILE; ++t) {
    const int x = tx * TILE + t % TILE, y = ty * TILE + t / TILE;
    if (!(x < a.width && y < a.height)) continue;
    const size_t pid = (size_t)a.width * y + x;
    const Px& p = px[t];
    s.final_T[pid] = p.T; s.n_contrib[pid] = p.last; s.wz[pid] = p.DD_D; s.wz2[pid] = p.DD_D2;
    for (int c = 0; c < 3; ++c) a.out_color[c * HW + pid] = p.C[c] + p.T * bgv(a.background, a.bg_mode, c, HW, pid);
    for (int c = 0; c < 7; ++c) a.out_phasor[c * HW + pid] = p.Pq[c] + p.T * bgv(a.background, a.bg_mode, c, HW, pid);
    a.out_depth[pid] = p.D; a.out_acc[pid] = p.A; a.out_depth_distortion[pid] = p.DD;
    for (int c = 0; c < 3; ++c) a.out_distribution[c * HW + pid] = p.WD[c];
  }
}

// per-Gaussian partial gradients of the blend, accumulated in double (the reference uses float
// atomics in arbitrary order; double accumulation is closer to the exact sum than either)
struct GAcc { double v[18]; };  // m2x m2y cx cy cw opac col[3] dist ndc ph[7]

void blend_tile_bwd(const GftBackwardArgs& a, const State& s, int tile, GAcc* acc) {
  const int tx = tile % s.gx, ty = tile / s.gx;
  const uint32_t r0 = s.ranges[2 * tile], r1 = s.ranges[2 * tile + 1];
  const int n = (int)(r1 - r0);
  const size_t HW = (size_t)a.width * a.height;
  const float ddelx_dx = 0.5f * a.width, ddely_dy = 0.5f * a.height;
  for (int t = 0; t < TILE * TILE; ++t) {
    const int x = tx * TILE + t % TILE, y = ty * TILE + t / TILE;
    if (!(x < a.width && y < a.height)) continue;
    const size_t pid = (size_t)a.width * y + x;
    const float pfx = (float)x, pfy = (float)y;
    const float T_final = s.final_T[pid];
    float T = T_final;
    const int last = (int)s.n_contrib[pid];
    const float wz_tot = s.wz[pid], wz2_tot = s.wz2[pid];
    float gC[3], gP[7];
    for (int c = 0; c < 3; ++c) gC[c] = a.dL_dout_color[c * HW + pid];
    for (int c = 0; c < 7; ++c) gP[c] = a.dL_dout_phasor[c * HW + pid];
    const float gD = a.dL_dout_depth[pid], gA = a.dL_dout_acc[pid], gDD = a.dL_dout_depth_distortion[pid];
    float bgc = 0, bgp = 0;
    for (int c = 0; c < 3; ++c) bgc += bgv(a.background, a.bg_mode, c, HW, pid) * gC[c];
    for (int c = 0; c < 7; ++c) bgp += bgv(a.background, a.bg_mode, c, HW, pid) * gP[c];
    float arC[3] = {0, 0, 0}, arP[7] = {0, 0, 0, 0, 0, 0, 0}, arD = 0, arA = 0, arDD = 0;
    float lastC[3] = {0, 0, 0}, lastP[7] = {0, 0, 0, 0, 0, 0, 0}, lastDist = 0, lastDLdw = 0, last_alpha = 0;
    for (int pos = std::min(n, last) - 1; pos >= 0; --pos) {  // backward.cu:735-741
      const uint32_t g = s.point_list[r0 + pos];
      const float* co = &s.conic_opacity[4 * (size_t)g];
      const float dx = s.means2D[2 * (size_t)g] - pfx, dy = s.means2D[2 * (size_t)g + 1] - pfy;
      const float power = pair_power(dx, dy, co[0], co[1], co[2]);
      if (power > 0.0f) continue;
      const float G = std::exp(power);
      const float alpha = std::fmin(0.99f, co[3] * G);
      if (alpha < 1.0f / 255.0f) continue;
      T = T / (1.f - alpha);
      const float w = alpha * T, wp = alpha * T * T;
      GAcc& A = acc[g];
      float dLa_c = 0, dLa_p = 0, dLa_d = 0, dLa_a = 0, dLa_dd = 0, dLa = 0;
      const float* rgb = &s.rgb[3 * (size_t)g];
      const float* ria = &s.ria[7 * (size_t)g];
      for (int c = 0; c < 3; ++c) {
        arC[c] = last_alpha * lastC[c] + (1.f - last_alpha) * arC[c];
        lastC[c] = rgb[c];
        dLa_c += (rgb[c] - arC[c]) * gC[c];
        A.v[6 + c] += (double)(w * gC[c]);
      }
      dLa_c *= T;
      for (int c = 0; c < 7; ++c) {
        arP[c] = last_alpha * lastP[c] + (1.f - last_alpha) * (1.f - last_alpha) * arP[c];
        lastP[c] = ria[c];
        dLa_p += (ria[c] - 2.f * (1.f - alpha) * arP[c]) * gP[c];
        A.v[11 + c] += (double)(wp * gP[c]);
      }
      dLa_p *= T * T;
      const float dist = s.dists[g];
      arD = last_alpha * lastDist + (1.f - last_alpha) * arD;
      lastDist = dist;
      dLa_d += (dist - arD) * gD;
      A.v[9] += (double)(w * gD);
      dLa_d *= T;
      arA = last_alpha + (1.f - last_alpha) * arA;
      dLa_a += (1.f - arA) * gA;
      dLa_a *= T;
      const float z = s.ndc[g];
      const float dL_dw = gDD * (z * z * (1 - T_final) - 2.0f * z * wz_tot + wz2_tot);
      arDD = last_alpha * lastDLdw + (1.f - last_alpha) * arDD;
      lastDLdw = dL_dw;
      dLa_dd += dL_dw - arDD;
      A.v[10] += (double)(gDD * 2.0f * alpha * T * (z * (1 - T_final) - wz_tot));
      dLa_dd *= T;
      last_alpha = alpha;
      dLa += (-T_final / (1.f - alpha)) * bgc;
      dLa_p += (-T_final / (1.f - alpha)) * bgp;
      dLa += dLa_c; dLa += dLa_p; dLa += dLa_d; dLa += dLa_a; dLa += dLa_dd;
      const float dL_dG = co[3] * dLa;
      const float gdx = G * dx, gdy = G * dy;
      const float dG_ddelx = -gdx * co[0] - gdy * co[1];
      const float dG_ddely = -gdy * co[2] - gdx * co[1];
      A.v[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
      A.v[1] += (double)(dL_dG * dG_ddely * ddely_dy);
      A.v[2] += (double)(-0.5f * gdx * dx * dL_dG);
      A.v[3] += (double)(-0.5f * gdx * dy * dL_dG);
      A.v[4] += (double)(-0.5f * gdy * dy * dL_dG);
      A.v[5] += (double)(G * dLa);
    }
  }
}

struct V3 { float x, y, z; };
V3 dnormvdv(V3 v, V3 dv) {  // auxiliary.h:114-124
  const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  const float inv = 1.0f / std::sqrt(sum2 * sum2 * sum2);
  return {((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * inv,
          (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * inv,
          (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * inv};
}

// SH backward for NC interleaved channels (backward.cu:20-139 with NC=3, :143-260 with NC=2)
template <int NC>
V3 sh_bwd(int deg, float x, float y, float z, const float* sh, const float* g, float* dsh) {
  float dX[NC] = {}, dY[NC] = {}, dZ[NC] = {};
  auto put = [&](int k, float b) { for (int c = 0; c < NC; ++c) dsh[k * NC + c] = b * g[c]; };
  put(0, C0);
  if (deg > 0) {
    put(1, -C1 * y); put(2, C1 * z); put(3, -C1 * x);
    for (int c = 0; c < NC; ++c) { dX[c] = -C1 * sh[3 * NC + c]; dY[c] = -C1 * sh[1 * NC + c]; dZ[c] = C1 * sh[2 * NC + c]; }
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      put(4, C2[0] * xy); put(5, C2[1] * yz); put(6, C2[2] * (2.f * zz - xx - yy));
      put(7, C2[3] * xz); put(8, C2[4] * (xx - yy));
      for (int c = 0; c < NC; ++c) {
        const float s4 = sh[4 * NC + c], s5 = sh[5 * NC + c], s6 = sh[6 * NC + c], s7 = sh[7 * NC + c], s8 = sh[8 * NC + c];
        dX[c] += C2[0] * y * s4 + C2[2] * 2.f * -x * s6 + C2[3] * z * s7 + C2[4] * 2.f * x * s8;
        dY[c] += C2[0] * x * s4 + C2[1] * z * s5 + C2[2] * 2.f * -y * s6 + C2[4] * 2.f * -y * s8;
        dZ[c] += C2[1] * y * s5 + C2[2] * 2.f * 2.f * z * s6 + C2[3] * x * s7;
      }
      if (deg > 2) {
        put(9, C3[0] * y * (3.f * xx - yy)); put(10, C3[1] * xy * z);
        put(11, C3[2] * y * (4.f * zz - xx - yy));
        put(12, C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
        put(13, C3[4] * x * (4.f * zz - xx - yy)); put(14, C3[5] * z * (xx - yy));
        put(15, C3[6] * x * (xx - 3.f * yy));
        for (int c = 0; c < NC; ++c) {
          const float s9 = sh[9 * NC + c], s10 = sh[10 * NC + c], s11 = sh[11 * NC + c], s12 = sh[12 * NC + c],
                      s13 = sh[13 * NC + c], s14 = sh[14 * NC + c], s15 = sh[15 * NC + c];
          dX[c] += (C3[0] * s9 * 3.f * 2.f * xy + C3[1] * s10 * yz + C3[2] * s11 * -2.f * xy +
                    C3[3] * s12 * -3.f * 2.f * xz + C3[4] * s13 * (-3.f * xx + 4.f * zz - yy) +
                    C3[5] * s14 * 2.f * xz + C3[6] * s15 * 3.f * (xx - yy));
          dY[c] += (C3[0] * s9 * 3.f * (xx - yy) + C3[1] * s10 * xz + C3[2] * s11 * (-3.f * yy + 4.f * zz - xx) +
                    C3[3] * s12 * -3.f * 2.f * yz + C3[4] * s13 * -2.f * xy + C3[5] * s14 * -2.f * yz +
                    C3[6] * s15 * -3.f * 2.f * xy);
          dZ[c] += (C3[1] * s10 * xy + C3[2] * s11 * 4.f * 2.f * yz + C3[3] * s12 * 3.f * (2.f * zz - xx - yy) +
                    C3[4] * s13 * 4.f * 2.f * xz + C3[5] * s14 * (xx - yy));
        }
      }
    }
  }
  V3 d = {0, 0, 0};
  for (int c = 0; c < NC; ++c) { d.x += dX[c] * g[c]; d.y += dY[c] * g[c]; d.z += dZ[c] * g[c]; }
  return d;
}

void preprocess_bwd_one(const GftBackwardArgs& a, const State& s, int i, const GAcc& A, float fx,
                        float fy, float d2p, double& sum_phase, double& sum_dc) {
  const float* V = a.viewmatrix;
  const float* pr = a.projmatrix;
  const float mx = a.means3D[3 * (size_t)i], my = a.means3D[3 * (size_t)i + 1], mz = a.means3D[3 * (size_t)i + 2];
  const float dm2x = (float)A.v[0], dm2y = (float)A.v[1];
  const float dcx = (float)A.v[2], dcy = (float)A.v[3], dcw = (float)A.v[4];
  float dcol[3] = {(float)A.v[6], (float)A.v[7], (float)A.v[8]};
  const float ddist = (float)A.v[9], dndc = (float)A.v[10];
  float dph[7];
  for (int c = 0; c < 7; ++c) dph[c] = (float)A.v[11 + c];
  a.dL_dmeans2D[3 * (size_t)i] = dm2x; a.dL_dmeans2D[3 * (size_t)i + 1] = dm2y; a.dL_dmeans2D[3 * (size_t)i + 2] = 0.f;
  a.dL_dopacity[i] = (float)A.v[5];
  if (a.dL_dcolors) for (int c = 0; c < 3; ++c) a.dL_dcolors[3 * (size_t)i + c] = dcol[c];
  if (a.dL_dphasors) for (int c = 0; c < 7; ++c) a.dL_dphasors[7 * (size_t)i + c] = dph[c];
  if (a.dL_dconic) { float* d = a.dL_dconic + 4 * (size_t)i; d[0] = dcx; d[1] = dcy; d[2] = 0; d[3] = dcw; }
  if (a.dL_ddist) a.dL_ddist[i] = ddist;
  if (a.dL_dndc) a.dL_dndc[i] = dndc;

  // ---- cov2D backward, backward.cu:276-394
  const float* v = a.cov3D_precomp ? a.cov3D_precomp + 6 * (size_t)i : &s.cov3D[6 * (size_t)i];
  float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
  float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
  const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
  const float mvx = tx, mvy = ty, mvz = tz;
  const float limx = 1.3f * a.tan_fovx, limy = 1.3f * a.tan_fovy;
  const float txtz = tx / tz, tytz = ty / tz;
  tx = std::fmin(limx, std::fmax(-limx, txtz)) * tz;
  ty = std::fmin(limy, std::fmax(-limy, tytz)) * tz;
  const float xgm = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  const float ygm = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  const float J00 = fx / tz, J02 = -(fx * tx) / (tz * tz), J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
  const float T00 = V[0] * J00 + V[2] * J02, T01 = V[4] * J00 + V[6] * J02, T02 = V[8] * J00 + V[10] * J02;
  const float T10 = V[1] * J11 + V[2] * J12, T11 = V[5] * J11 + V[6] * J12, T12 = V[9] * J11 + V[10] * J12;
  const float B00 = T00 * v[0] + T01 * v[1] + T02 * v[2], B01 = T00 * v[1] + T01 * v[3] + T02 * v[4],
              B02 = T00 * v[2] + T01 * v[4] + T02 * v[5];
  const float B10 = T10 * v[0] + T11 * v[1] + T12 * v[2], B11 = T10 * v[1] + T11 * v[3] + T12 * v[4],
              B12 = T10 * v[2] + T11 * v[4] + T12 * v[5];
  const float ca = T00 * B00 + T01 * B01 + T02 * B02 + 0.3f;
  const float cb = T00 * B10 + T01 * B11 + T02 * B12;
  const float cc = T10 * B10 + T11 * B11 + T12 * B12 + 0.3f;
  const float denom = ca * cc - cb * cb;
  float da = 0, db = 0, dc_ = 0;
  const float d2inv = 1.0f / ((denom * denom) + 0.0000001f);
  float dcov[6] = {0, 0, 0, 0, 0, 0};
  if (d2inv != 0) {
    da = d2inv * (-cc * cc * dcx + 2 * cb * cc * dcy + (denom - ca * cc) * dcw);
    dc_ = d2inv * (-ca * ca * dcw + 2 * ca * cb * dcy + (denom - ca * cc) * dcx);
    db = d2inv * 2 * (cb * cc * dcx - (denom + 2 * cb * cb) * dcy + ca * cb * dcw);
    dcov[0] = (T00 * T00 * da + T00 * T10 * db + T10 * T10 * dc_);
    dcov[3] = (T01 * T01 * da + T01 * T11 * db + T11 * T11 * dc_);
    dcov[5] = (T02 * T02 * da + T02 * T12 * db + T12 * T12 * dc_);
    dcov[1] = 2 * T00 * T01 * da + (T00 * T11 + T01 * T10) * db + 2 * T10 * T11 * dc_;
    dcov[2] = 2 * T00 * T02 * da + (T00 * T12 + T02 * T10) * db + 2 * T10 * T12 * dc_;
    dcov[4] = 2 * T02 * T01 * da + (T01 * T12 + T02 * T11) * db + 2 * T11 * T12 * dc_;
  }
  if (a.dL_dcov3D) for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * (size_t)i + k] = dcov[k];
  const float dT00 = 2 * B00 * da + B10 * db, dT01 = 2 * B01 * da + B11 * db, dT02 = 2 * B02 * da + B12 * db;
  const float dT10 = 2 * B10 * dc_ + B00 * db, dT11 = 2 * B11 * dc_ + B01 * db, dT12 = 2 * B12 * dc_ + B02 * db;
  const float dJ00 = V[0] * dT00 + V[4] * dT01 + V[8] * dT02;
  const float dJ02 = V[2] * dT00 + V[6] * dT01 + V[10] * dT02;
  const float dJ11 = V[1] * dT10 + V[5] * dT11 + V[9] * dT12;
  const float dJ12 = V[2] * dT10 + V[6] * dT11 + V[10] * dT12;
  const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
  const float dtx = xgm * -fx * itz2 * dJ02, dty = ygm * -fy * itz2 * dJ12;
  const float dtz = -fx * itz2 * dJ00 - fy * itz2 * dJ11 + (2 * fx * tx) * itz3 * dJ02 + (2 * fy * ty) * itz3 * dJ12;
  float gmx = V[0] * dtx + V[1] * dty + V[2] * dtz;
  float gmy = V[4] * dtx + V[5] * dty + V[6] * dtz;
  float gmz = V[8] * dtx + V[9] * dty + V[10] * dtz;

  // ---- mean2D -> mean3D, backward.cu:498-519
  {
    const float mw = 1.0f / ((pr[3] * mx + pr[7] * my + pr[11] * mz + pr[15]) + 0.0000001f);
    const float mul1 = (pr[0] * mx + pr[4] * my + pr[8] * mz + pr[12]) * mw * mw;
    const float mul2 = (pr[1] * mx + pr[5] * my + pr[9] * mz + pr[13]) * mw * mw;
    gmx += (pr[0] * mw - pr[3] * mul1) * dm2x + (pr[1] * mw - pr[3] * mul2) * dm2y;
    gmy += (pr[4] * mw - pr[7] * mul1) * dm2x + (pr[5] * mw - pr[7] * mul2) * dm2y;
    gmz += (pr[8] * mw - pr[11] * mul1) * dm2x + (pr[9] * mw - pr[11] * mul2) * dm2y;
  }
  V3 dorig = {0, 0, 0};
  float ux = 0, uy = 0, uz = 0;
  if (a.shs || a.shs_p) {
    dorig = {mx - a.campos[0], my - a.campos[1], mz - a.campos[2]};
    const float len = std::sqrt(dorig.x * dorig.x + dorig.y * dorig.y + dorig.z * dorig.z);
    ux = dorig.x / len; uy = dorig.y / len; uz = dorig.z / len;
  }
  const int nco = (a.sh_degree + 1) * (a.sh_degree + 1);
  if (a.shs) {
    float g[3] = {dcol[0], dcol[1], dcol[2]};
    for (int c = 0; c < 3; ++c) g[c] *= s.clamped[3 * (size_t)i + c] ? 0.f : 1.f;
    float* dsh = a.dL_dsh + (size_t)i * a.M * 3;
    const V3 dd = sh_bwd<3>(a.sh_degree, ux, uy, uz, a.shs + (size_t)i * a.M * 3, g, dsh);
    for (int k = 3 * nco; k < 3 * a.M; ++k) dsh[k] = 0.f;
    const V3 dm = dnormvdv(dorig, dd);
    gmx += dm.x; gmy += dm.y; gmz += dm.z;
  }
  const float dist = s.dists[i];
  if (a.shs_p) {  // backward.cu:527-587
    float phase = dist * d2p + a.phase_offset;
    if (a.use_view_dependent_phase) phase += s.pa[2 * (size_t)i];
    const float amp = s.pa[2 * (size_t)i + 1];
    const float factor = 1.0f / (dist * dist);
    const float dR = dph[0], dI = dph[1], dA = dph[2], q1 = dph[3], q2 = dph[4], q3 = dph[5], q4 = dph[6];
    const float sp = std::sin(phase), cp = std::cos(phase), dc = a.dc_offset;
    const float dsum = dR * -sp + dI * cp + q1 * -sp + q2 * sp + q3 * cp + q4 * -cp;
    float gpa[2] = {0, 0};
    if (a.use_view_dependent_phase) gpa[0] = dsum * amp * factor;
    sum_phase += (double)(dsum * amp * factor);
    gpa[1] = (dR * cp + dI * sp + dA + q1 * (cp + dc) + q2 * (-cp + dc) + q3 * (sp + dc) + q4 * (-sp + dc)) * factor;
    sum_dc += (double)((q1 + q2 + q3 + q4) * amp * factor);
    const float coeff = dsum * d2p * amp * factor / dist +
                        (dR * -cp + dI * -sp - dA + q1 * -(cp + dc) + q2 * (cp - dc) + q3 * -(sp + dc) + q4 * (sp - dc)) *
                            2.0f * amp * factor * factor;
    const float xv = mvx * coeff, yv = mvy * coeff, zv = mvz * coeff;
    gmx += xv * V[0] + yv * V[1] + zv * V[2];
    gmy += xv * V[4] + yv * V[5] + zv * V[6];
    gmz += xv * V[8] + yv * V[9] + zv * V[10];
    gpa[1] *= s.clamped_p[i] ? 0.f : 1.f;
    float* dshp = a.dL_dsh_p + (size_t)i * a.M_p * 2;
    const V3 dd = sh_bwd<2>(a.sh_degree, ux, uy, uz, a.shs_p + (size_t)i * a.M_p * 2, gpa, dshp);
    for (int k = 2 * nco; k < 2 * a.M_p; ++k) dshp[k] = 0.f;
    const V3 dm = dnormvdv(dorig, dd);
    gmx += dm.x; gmy += dm.y; gmz += dm.z;
  }
  {  // backward.cu:589-601
    const float dn = (a.far_n * a.near_n) / ((a.far_n - a.near_n) * dist * dist);
    const float dL = dndc * dn + ddist;
    const float xv = dL * mvx / dist, yv = dL * mvy / dist, zv = dL * mvz / dist;
    gmx += xv * V[0] + yv * V[1] + zv * V[2];
    gmy += xv * V[4] + yv * V[5] + zv * V[6];
    gmz += xv * V[8] + yv * V[9] + zv * V[10];
  }
  a.dL_dmeans3D[3 * (size_t)i] = gmx; a.dL_dmeans3D[3 * (size_t)i + 1] = gmy; a.dL_dmeans3D[3 * (size_t)i + 2] = gmz;

  if (a.scales) {  // backward.cu:399-462
    const float* q = a.rotations + 4 * (size_t)i;
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float sx = a.scale_modifier * a.scales[3 * (size_t)i], sy = a.scale_modifier * a.scales[3 * (size_t)i + 1],
                sz = a.scale_modifier * a.scales[3 * (size_t)i + 2];
    const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
    const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
    const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
    const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02, M10 = sx * R10, M11 = sy * R11, M12 = sz * R12,
                M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
    const float S00 = dcov[0], S01 = 0.5f * dcov[1], S02 = 0.5f * dcov[2], S11 = dcov[3], S12 = 0.5f * dcov[4], S22 = dcov[5];
    const float D00 = 2.f * (M00 * S00 + M10 * S01 + M20 * S02), D01 = 2.f * (M01 * S00 + M11 * S01 + M21 * S02),
                D02 = 2.f * (M02 * S00 + M12 * S01 + M22 * S02);
    const float D10 = 2.f * (M00 * S01 + M10 * S11 + M20 * S12), D11 = 2.f * (M01 * S01 + M11 * S11 + M21 * S12),
                D12 = 2.f * (M02 * S01 + M12 * S11 + M22 * S12);
    const float D20 = 2.f * (M00 * S02 + M10 * S12 + M20 * S22), D21 = 2.f * (M01 * S02 + M11 * S12 + M21 * S22),
                D22 = 2.f * (M02 * S02 + M12 * S12 + M22 * S22);
    a.dL_dscales[3 * (size_t)i] = R00 * D00 + R10 * D10 + R20 * D20;
    a.dL_dscales[3 * (size_t)i + 1] = R01 * D01 + R11 * D11 + R21 * D21;
    a.dL_dscales[3 * (size_t)i + 2] = R02 * D02 + R12 * D12 + R22 * D22;
    const float t00 = D00 * sx, t01 = D10 * sx, t02 = D20 * sx, t10 = D01 * sy, t11 = D11 * sy, t12 = D21 * sy,
                t20 = D02 * sz, t21 = D12 * sz, t22 = D22 * sz;
    float* dq = a.dL_drotations + 4 * (size_t)i;
    dq[0] = 2 * z * (t01 - t10) + 2 * y * (t20 - t02) + 2 * x * (t12 - t21);
    dq[1] = 2 * y * (t10 + t01) + 2 * z * (t20 + t02) + 2 * r * (t12 - t21) - 4 * x * (t22 + t11);
    dq[2] = 2 * x * (t10 + t01) + 2 * r * (t20 - t02) + 2 * z * (t12 + t21) - 4 * y * (t22 + t00);
    dq[3] = 2 * r * (t01 - t10) + 2 * x * (t20 + t02) + 2 * y * (t12 + t21) - 4 * z * (t11 + t00);
  }
}

}  // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }
void orc_free(void* h) { delete reinterpret_cast<State*>(h); }

// Forward on HOST pointers.  Outputs must be caller-allocated; they are fully written (zeros where
// the reference leaves its zero fill).  *handle receives the saved state for orc_backward.
int orc_forward(const GftForwardArgs* ap, void** handle) {
  const GftForwardArgs& a = *ap;
  State* sp = new State();
  State& s = *sp;
  *handle = sp;
  const int P = a.P, W = a.width, H = a.height;
  s.P = P; s.W = W; s.H = H;
  s.gx = (W + TILE - 1) / TILE; s.gy = (H + TILE - 1) / TILE;
  const int T = s.gx * s.gy;
  const size_t N = (size_t)W * H;
  std::fill(a.out_color, a.out_color + 3 * N, 0.f);
  std::fill(a.out_phasor, a.out_phasor + 7 * N, 0.f);
  std::fill(a.out_depth, a.out_depth + N, 0.f);
  std::fill(a.out_acc, a.out_acc + N, 0.f);
  std::fill(a.out_depth_distortion, a.out_depth_distortion + N, 0.f);
  std::fill(a.out_distribution, a.out_distribution + 3 * N, 0.f);
  if (a.out_normal) std::fill(a.out_normal, a.out_normal + 3 * N, 0.f);
  if (a.out_entropy) std::fill(a.out_entropy, a.out_entropy + N, 0.f);
  if (a.out_amp_distortion) std::fill(a.out_amp_distortion, a.out_amp_distortion + N, 0.f);
  s.final_T.assign(N, 0.f); s.wz.assign(N, 0.f); s.wz2.assign(N, 0.f); s.n_contrib.assign(N, 0u);
  s.ranges.assign(2 * (size_t)T, 0u);
  if (P == 0) return 0;  // rasterize_points.cu:104
  s.depths.assign(P, 0.f); s.ndc.assign(P, 0.f); s.dists.assign(P, 0.f);
  s.means2D.assign(2 * (size_t)P, 0.f); s.cov3D.assign(6 * (size_t)P, 0.f);
  s.conic_opacity.assign(4 * (size_t)P, 0.f); s.rgb.assign(3 * (size_t)P, 0.f);
  s.ria.assign(7 * (size_t)P, 0.f); s.pa.assign(2 * (size_t)P, 0.f);
  s.clamped.assign(3 * (size_t)P, 0); s.clamped_p.assign(P, 0);
  s.tiles_touched.assign(P, 0u); s.point_offsets.assign(P, 0u); s.rect.assign(4 * (size_t)P, 0u);
  const float fy = H / (2.0f * a.tan_fovy), fx = W / (2.0f * a.tan_fovx);  // rasterizer_impl.cu:249-250
  const float d2p = 4.0f * PI_F / a.depth_range;                          // forward.cu:752
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; ++i) preprocess_one(a, s, i, fx, fy, d2p);
  uint64_t run = 0;  // rasterizer_impl.cu:307
  for (int i = 0; i < P; ++i) { run += s.tiles_touched[i]; s.point_offsets[i] = (uint32_t)run; }
  if (run > 0x7fffffffull) { g_err = "num_rendered exceeds 2^31-1"; return -4; }
  const int R = (int)run;
  s.R = R;
  s.keys.resize(R); s.point_list.resize(R);
  std::vector<uint64_t> ukeys(R);
  std::vector<uint32_t> uvals(R);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; ++i) {  // rasterizer_impl.cu:72-113
    if (a.radii[i] <= 0) continue;
    uint32_t off = i == 0 ? 0u : s.point_offsets[i - 1];
    const uint32_t* rc = &s.rect[4 * (size_t)i];
    uint32_t dbits;
    std::memcpy(&dbits, &s.depths[i], 4);
    for (uint32_t y = rc[1]; y < rc[3]; ++y)
      for (uint32_t x = rc[0]; x < rc[2]; ++x) {
        ukeys[off] = ((uint64_t)(y * (uint32_t)s.gx + x) << 32) | dbits;
        uvals[off] = (uint32_t)i;
        ++off;
      }
  }
  // stable sort on bits [0, 32 + tile_bits) — all significant bits, rasterizer_impl.cu:331-339
  {
    std::vector<uint32_t> order(R);
    std::iota(order.begin(), order.end(), 0u);
    const int bits = 32 + tile_bits((uint32_t)T);
    const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t l, uint32_t r) { return (ukeys[l] & mask) < (ukeys[r] & mask); });
    for (int k = 0; k < R; ++k) { s.keys[k] = ukeys[order[k]]; s.point_list[k] = uvals[order[k]]; }
  }
  for (int k = 0; k < R; ++k) {  // rasterizer_impl.cu:118-140
    const uint32_t cur = (uint32_t)(s.keys[k] >> 32);
    if (k == 0) s.ranges[2 * cur] = 0;
    else {
      const uint32_t prev = (uint32_t)(s.keys[k - 1] >> 32);
      if (cur != prev) { s.ranges[2 * prev + 1] = k; s.ranges[2 * cur] = k; }
    }
    if (k == R - 1) s.ranges[2 * cur + 1] = R;
  }
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<std::vector<float>> counts(nthreads, std::vector<float>(P, 0.f));
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < T; ++t) {
    int th = 0;
#ifdef _OPENMP
    th = omp_get_thread_num();
#endif
    blend_tile_fwd(a, s, t, counts[th]);
  }
  for (int th = 0; th < nthreads; ++th)
    for (int i = 0; i < P; ++i) a.pixels[i] += counts[th][i];  // integer-valued, exact in any order
  return R;
}

int orc_backward(const GftBackwardArgs* ap, void* handle) {
  const GftBackwardArgs& a = *ap;
  const State& s = *reinterpret_cast<State*>(handle);
  const int P = a.P;
  if (a.dL_dphase_offset) *a.dL_dphase_offset = 0.f;
  if (a.dL_ddc_offset) *a.dL_ddc_offset = 0.f;
  if (P == 0) return 0;
  const int T = s.gx * s.gy;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<std::vector<GAcc>> accs(nthreads);
#pragma omp parallel
  {
    int th = 0;
#ifdef _OPENMP
    th = omp_get_thread_num();
#endif
    accs[th].assign(P, GAcc{});
#pragma omp for schedule(dynamic, 1)
    for (int t = 0; t < T; ++t) blend_tile_bwd(a, s, t, accs[th].data());
  }
  std::vector<GAcc>& acc = accs[0];
  for (int th = 1; th < nthreads; ++th)
    for (int i = 0; i < P; ++i)
      for (int k = 0; k < 18; ++k) acc[i].v[k] += accs[th][i].v[k];
  const float fy = a.height / (2.0f * a.tan_fovy), fx = a.width / (2.0f * a.tan_fovx);
  const float d2p = 4.0f * PI_F / a.depth_range;  // backward.cu:936
  double sp = 0, sd = 0;
#pragma omp parallel for schedule(static) reduction(+ : sp, sd)
  for (int i = 0; i < P; ++i) {
    if (a.radii[i] > 0) {
      preprocess_bwd_one(a, s, i, acc[i], fx, fy, d2p, sp, sd);
    } else {  // untouched zero-filled rows in the reference
      for (int k = 0; k < 3; ++k) { a.dL_dmeans2D[3 * (size_t)i + k] = 0; a.dL_dmeans3D[3 * (size_t)i + k] = 0; }
      a.dL_dopacity[i] = 0;
      if (a.dL_dsh) std::fill(a.dL_dsh + (size_t)i * a.M * 3, a.dL_dsh + (size_t)(i + 1) * a.M * 3, 0.f);
      if (a.dL_dsh_p) std::fill(a.dL_dsh_p + (size_t)i * a.M_p * 2, a.dL_dsh_p + (size_t)(i + 1) * a.M_p * 2, 0.f);
      if (a.dL_dscales) for (int k = 0; k < 3; ++k) a.dL_dscales[3 * (size_t)i + k] = 0;
      if (a.dL_drotations) for (int k = 0; k < 4; ++k) a.dL_drotations[4 * (size_t)i + k] = 0;
      if (a.dL_dcolors) for (int k = 0; k < 3; ++k) a.dL_dcolors[3 * (size_t)i + k] = 0;
      if (a.dL_dphasors) for (int k = 0; k < 7; ++k) a.dL_dphasors[7 * (size_t)i + k] = 0;
      if (a.dL_dcov3D) for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * (size_t)i + k] = 0;
      if (a.dL_dconic) for (int k = 0; k < 4; ++k) a.dL_dconic[4 * (size_t)i + k] = 0;
      if (a.dL_ddist) a.dL_ddist[i] = 0;
      if (a.dL_dndc) a.dL_dndc[i] = 0;
    }
  }
  if (a.dL_dphase_offset) *a.dL_dphase_offset = (float)sp;
  if (a.dL_ddc_offset) *a.dL_ddc_offset = (float)sd;
  return 0;
}

// Saved-state accessor for the tests: name -> pointer + element count.
int orc_get(void* handle, const char* name, const void** ptr, size_t* count) {
  State& s = *reinterpret_cast<State*>(handle);
#define ORC_FIELD(n, vec) if (!std::strcmp(name, n)) { *ptr = (vec).data(); *count = (vec).size(); return 0; }
  ORC_FIELD("depths", s.depths) ORC_FIELD("ndc", s.ndc) ORC_FIELD("dists", s.dists)
  ORC_FIELD("means2D", s.means2D) ORC_FIELD("cov3D", s.cov3D) ORC_FIELD("conic_opacity", s.conic_opacity)
  ORC_FIELD("rgb", s.rgb) ORC_FIELD("real_img_amp", s.ria) ORC_FIELD("pa", s.pa)
  ORC_FIELD("clamped", s.clamped) ORC_FIELD("clamped_p", s.clamped_p)
  ORC_FIELD("tiles_touched", s.tiles_touched) ORC_FIELD("point_offsets", s.point_offsets)
  ORC_FIELD("keys", s.keys) ORC_FIELD("point_list", s.point_list) ORC_FIELD("ranges", s.ranges)
  ORC_FIELD("final_T", s.final_T) ORC_FIELD("w_z_total", s.wz) ORC_FIELD("w_z2_total", s.wz2)
  ORC_FIELD("n_contrib", s.n_contrib)
#undef ORC_FIELD
  return -1;
}

int orc_mark_visible(int P, const float* means3D, const float* V, uint8_t* present, float near_n,
                     float far_n) {  // rasterizer_impl.cu:54-68
  for (int i = 0; i < P; ++i) {
    const float vz = xf(V, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    present[i] = !(vz < near_n || vz > far_n);
  }
  return 0;
}

// distCUDA2: mean of the three smallest squared distances to the OTHER points (coincident points
// count with distance 0), simple_knn.cu:131-183.  Brute force; squared distance in the reference
// binary's contracted form FMUL(dy,dy) -> FFMA(dx,dx,.) -> FFMA(dz,dz,.).
int orc_dist2(const float* pts, int P, float* out) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; ++i) {
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    for (int j = 0; j < P; ++j) {
      if (j == i) continue;
      const float dx = pts[3 * (size_t)j] - x, dy = pts[3 * (size_t)j + 1] - y, dz = pts[3 * (size_t)j + 2] - z;
      float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
      if (b0 > d) std::swap(b0, d);
      if (b1 > d) std::swap(b1, d);
      if (b2 > d) b2 = d;
    }
    out[i] = ((b0 + b1) + b2) / 3.0f;
  }
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
