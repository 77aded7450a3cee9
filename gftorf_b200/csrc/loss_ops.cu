// loss_ops.cu — photometric / ToF loss on a rendered image: value and d loss / d image in two
// launches, hand-written for sm_100a.  C ABI: include/gftorf_train.h.
//
// Replaces, per loss term of train.py:204-223: the elementwise chain of l1_loss / l2_loss /
// weighted_l1_loss / weighted_l{1,2}_loss_quad (utils/loss_utils.py:17-33), the five depthwise
// 11x11 conv2d of _ssim (:97-103) with their ~20 elementwise kernels, and everything autograd
// replays backwards through them (five more convolutions).
//
//   pass 1 (ssim_fwd_kernel)  per 16x16 tile and channel: stage img and gt with a 5-pixel zero halo,
//                             separable 11-tap Gaussian of (x, y, xx, yy, xy), SSIM map -> block
//                             sum; the three partials d map/d mu1, d map/d sigma1^2, d map/d sigma12
//                             go to scratch.
//   pass 2 (loss_bwd_kernel)  separable Gaussian of the three partial maps (zero outside the image,
//                             the adjoint of zero padding), combined with img/gt into the SSIM
//                             gradient, plus the elementwise term's value and gradient.
// With lambda_dssim == 0 only the elementwise part of pass 2 runs.
#include <cuda_runtime.h>
#include <cmath>

#include "../../include/gftorf_train.h"
#include "kernels.h"

namespace gft {
namespace {

constexpr int LT = 16;            // tile edge
constexpr int HALO = 5;           // window 11
constexpr int LS = LT + 2 * HALO; // 26

struct Window { float w[11]; };

__device__ __forceinline__ float block_reduce_add(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int t = threadIdx.y * LT + threadIdx.x;
  if ((t & 31) == 0) s_red[t >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (t < 32) {
    r = t < (LT * LT / 32) ? s_red[t] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;   // valid in thread 0
}

__global__ void __launch_bounds__(LT * LT)
ssim_fwd_kernel(GftLossArgs a, Window win, float* __restrict__ ssim_sum) {
  __shared__ float sx[LS][LS + 1], sy[LS][LS + 1];
  __shared__ float hz[5][LS][LT + 1];
  __shared__ float s_red[LT * LT / 32];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t plane = (size_t)a.H * a.W;
  const float* img = a.img + c * plane;
  const float* gt = a.gt + c * plane;
  const int t = threadIdx.y * LT + threadIdx.x;
  for (int e = t; e < LS * LS; e += LT * LT) {
    const int r = e / LS, q = e - r * LS;
    const int gy = y0 + r - HALO, gx = x0 + q - HALO;
    const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
    sx[r][q] = in ? __ldg(img + (size_t)gy * a.W + gx) : 0.f;
    sy[r][q] = in ? __ldg(gt + (size_t)gy * a.W + gx) : 0.f;
  }
  __syncthreads();
  for (int e = t; e < LS * LT; e += LT * LT) {       // horizontal pass: 26 rows x 16 columns
    const int r = e / LT, q = e - r * LT;
    float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float u = sx[r][q + k], v = sy[r][q + k], w = win.w[k];
      m1 += w * u; m2 += w * v; xx += w * (u * u); yy += w * (v * v); xy += w * (u * v);
    }
    hz[0][r][q] = m1; hz[1][r][q] = m2; hz[2][r][q] = xx; hz[3][r][q] = yy; hz[4][r][q] = xy;
  }
  __syncthreads();
  const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
  float val = 0.f;
  if (px < a.W && py < a.H) {
    float mu1 = 0.f, mu2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = win.w[k];
      mu1 += w * hz[0][threadIdx.y + k][threadIdx.x];
      mu2 += w * hz[1][threadIdx.y + k][threadIdx.x];
      xx += w * hz[2][threadIdx.y + k][threadIdx.x];
      yy += w * hz[3][threadIdx.y + k][threadIdx.x];
      xy += w * hz[4][threadIdx.y + k][threadIdx.x];
    }
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s1 = xx - mu1_sq, s2 = yy - mu2_sq, s12 = xy - mu12;
    const float A = 2.f * mu12 + C1, B = 2.f * s12 + C2, Cc = mu1_sq + mu2_sq + C1, D = s1 + s2 + C2;
    val = (A * B) / (Cc * D);
    // partial derivatives of the map w.r.t. the three window statistics that depend on img
    const float dmu1 = (mu2 * 2.f * B) / (Cc * D) - (mu2 * 2.f * A) / (Cc * D) -
                       (mu1 * 2.f * A * B) / (Cc * Cc * D) + (mu1 * 2.f * A * B) / (Cc * D * D);
    const float ds1 = -(A * B) / (Cc * D * D);
    const float ds12 = (2.f * A) / (Cc * D);
    const size_t n = (size_t)a.C * plane;
    const size_t o = c * plane + (size_t)py * a.W + px;
    a.scratch[o] = dmu1;
    a.scratch[n + o] = ds1;
    a.scratch[2 * n + o] = ds12;
  }
  const float tot = block_reduce_add(val, s_red);
  if (t == 0) atomicAdd(ssim_sum, tot);
}

__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

template <bool SSIM>
__global__ void __launch_bounds__(LT * LT)
loss_bwd_kernel(GftLossArgs a, Window win, const float* __restrict__ ssim_sum, int zero_sum_after) {
  __shared__ float sm[SSIM ? 3 : 1][SSIM ? LS : 1][LS + 1];
  __shared__ float hz[SSIM ? 3 : 1][SSIM ? LS : 1][LT + 1];
  __shared__ float s_red[LT * LT / 32];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t plane = (size_t)a.H * a.W;
  const size_t n = (size_t)a.C * plane;
  const int t = threadIdx.y * LT + threadIdx.x;
  if (SSIM) {
    for (int e = t; e < LS * LS; e += LT * LT) {
      const int r = e / LS, q = e - r * LS;
      const int gy = y0 + r - HALO, gx = x0 + q - HALO;
      const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
      const size_t o = c * plane + (size_t)gy * a.W + gx;
      sm[0][r][q] = in ? __ldg(a.scratch + o) : 0.f;
      sm[1][r][q] = in ? __ldg(a.scratch + n + o) : 0.f;
      sm[2][r][q] = in ? __ldg(a.scratch + 2 * n + o) : 0.f;
    }
    __syncthreads();
    for (int e = t; e < LS * LT; e += LT * LT) {
      const int r = e / LT, q = e - r * LT;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float w = win.w[k];
        v0 += w * sm[0][r][q + k]; v1 += w * sm[1][r][q + k]; v2 += w * sm[2][r][q + k];
      }
      hz[0][r][q] = v0; hz[1][r][q] = v1; hz[2][r][q] = v2;
    }
    __syncthreads();
  }
  const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
  float val = 0.f;
  if (px < a.W && py < a.H) {
    const size_t pix = (size_t)py * a.W + px;
    const float x = __ldg(a.img + c * plane + pix), y = __ldg(a.gt + c * plane + pix);
    const float d = x - y;
    const float inv_n = 1.0f / (float)n;
    float g = 0.f;
    switch (a.kind) {
      case 0: val = fabsf(d) * inv_n; g = sgn(d) * inv_n; break;
      case 1: val = d * d * inv_n; g = 2.f * d * inv_n; break;
      case 2: {
        if (c < a.nch) {
          float ss = 0.f;
          for (int k = 0; k < a.C; ++k) { const float u = __ldg(a.img + k * plane + pix); ss += u * u; }
          const float wgt = a.w + sqrtf(ss);
          const float inv = 1.0f / ((float)a.nch * (float)plane);
          val = fabsf(d / wgt) * inv; g = sgn(d) / wgt * inv;
        }
      } break;
      case 3: { const float wgt = a.w + fabsf(x); val = fabsf(d / wgt) * inv_n; g = sgn(d) / wgt * inv_n; } break;
      default: { const float wgt = a.w + fabsf(x); const float r = d / wgt; val = r * r * inv_n; g = 2.f * r / wgt * inv_n; } break;
    }
    g *= a.lambda * (1.f - a.lambda_dssim);
    val *= a.lambda * (1.f - a.lambda_dssim);
    if (SSIM) {
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float w = win.w[k];
        c0 += w * hz[0][threadIdx.y + k][threadIdx.x];
        c1 += w * hz[1][threadIdx.y + k][threadIdx.x];
        c2 += w * hz[2][threadIdx.y + k][threadIdx.x];
      }
      // d(sum of the map)/d img, times d loss/d map = -lambda * lambda_dssim / n
      g += -(a.lambda * a.lambda_dssim * inv_n) * (c0 + 2.f * x * c1 + y * c2);
    }
    a.grad[c * plane + pix] = g;
  }
  const float tot = block_reduce_add(val, s_red);
  if (t == 0) {
    float add = tot;
    // one block also books the SSIM term's value: lambda * lambda_dssim * (1 - mean(map))
    if (SSIM && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
      add += a.lambda * a.lambda_dssim * (1.f - __ldg(ssim_sum) / (float)n);
    atomicAdd(a.loss, add);
  }
  (void)zero_sum_after;
}

Window make_window() {
  // loss_utils.py:75-77: exp(-(x - 5)^2 / (2 * 1.5^2)) as float32, normalised by its float32 sum
  Window w;
  float s = 0.f;
  for (int i = 0; i < 11; ++i) {
    w.w[i] = (float)std::exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    s += w.w[i];
  }
  for (int i = 0; i < 11; ++i) w.w[i] /= s;
  return w;
}

}  // namespace
}  // namespace gft

extern "C" {

size_t gft_fused_loss_scratch_bytes(int C, int H, int W) {
  // three partial-derivative maps + the SSIM sum accumulator
  return ((size_t)3 * C * H * W + 64) * sizeof(float);
}

int gft_fused_loss(const GftLossArgs* a, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return gft::set_error(-1, "gft_fused_loss: null args");
  if (a->C <= 0 || a->H <= 0 || a->W <= 0) return gft::set_error(-1, "gft_fused_loss: bad shape");
  if (a->kind < 0 || a->kind > 4) return gft::set_error(-1, "gft_fused_loss: kind must be 0..4");
  if (a->kind == 2 && (a->nch <= 0 || a->nch > a->C)) return gft::set_error(-1, "gft_fused_loss: bad nch");
  if (!a->img || !a->gt || !a->grad || !a->loss) return gft::set_error(-1, "gft_fused_loss: null pointer");
  const bool ssim = a->lambda_dssim != 0.f;
  if (ssim && !a->scratch) return gft::set_error(-1, "gft_fused_loss: scratch is required for SSIM");
  const gft::Window win = gft::make_window();
  const dim3 grid((a->W + gft::LT - 1) / gft::LT, (a->H + gft::LT - 1) / gft::LT, a->C);
  const dim3 block(gft::LT, gft::LT);
  if (ssim) {
    float* sum = a->scratch + (size_t)3 * a->C * a->H * a->W;
    cudaMemsetAsync(sum, 0, sizeof(float), stream);
    gft::ssim_fwd_kernel<<<grid, block, 0, stream>>>(*a, win, sum);
    gft::loss_bwd_kernel<true><<<grid, block, 0, stream>>>(*a, win, sum, 0);
    gft::note_launches(2);
  } else {
    gft::loss_bwd_kernel<false><<<grid, block, 0, stream>>>(*a, win, nullptr, 0);
    gft::note_launches(1);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
