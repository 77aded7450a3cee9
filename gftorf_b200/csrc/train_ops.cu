// train_ops.cu — the streaming per-Gaussian operators either side of the rasterizer in a training
// iteration, hand-written for sm_100a.  C ABI: include/gftorf_train.h.
//
//   assemble_fwd_kernel   raw parameters -> rasterizer inputs   (gaussian_renderer/__init__.py:81-105,
//                         scene/gaussian_model.py:123-157): one launch instead of 7 zero fills,
//                         14 masked index-copies, sigmoid, exp, normalize and 3 torch.cat
//   assemble_bwd_kernel   gradients of the rasterizer inputs -> gradients of the raw parameters and
//                         of the deformation outputs (what autograd derives for that chain)
//   adam_kernel           every parameter group of scene/gaussian_model.py:247-272 in one pass
//
// All three are HBM-bound streams: 128-bit accesses where rows allow it, every output element
// written exactly once, no zero-fill passes.
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>

#include "../../include/gftorf_train.h"
#include "common.cuh"
#include "kernels.h"

namespace gft {

namespace {

constexpr int ASM_BLOCK = 256;

// One block = 256 consecutive Gaussians.  Phase A: one thread per Gaussian for the short rows
// (position, opacity, scale, rotation).  Phase B/C: the block's SH rows are contiguous chunks of
// the output (256 x 3M and 256 x 2M floats): all threads walk them element by element, so stores
// are fully coalesced and the gathers from the dc / rest / deformation arrays are contiguous runs.
// MC: compile-time SH coefficient count (16 = the reference's sh_degree 3, the row lengths become
// constants and the per-element divisions multiplications), 0 = take it from the arguments
template <int MC>
__global__ void __launch_bounds__(ASM_BLOCK) assemble_fwd_kernel(GftAssembleArgs a) {
  const int first = blockIdx.x * ASM_BLOCK;
  const int i = first + (int)threadIdx.x;
  const int M = MC ? MC : a.M;
  __shared__ int s_dyn[ASM_BLOCK];    // deformation row, -1: static, -2: excluded region / out of range
  int code = -2;
  if (i < a.P) {
    const int j = a.dyn_index ? __ldg(a.dyn_index + i) : -1;
    const bool on = j >= 0 ? (a.include_dynamic != 0) : (a.include_static != 0);
    code = on ? (j >= 0 ? j : -1) : -2;
    float3 m = make_float3(0.f, 0.f, 0.f), s = m;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    float o = 0.f;
    if (on) {
      m = make_float3(__ldg(a.xyz + 3 * (size_t)i), __ldg(a.xyz + 3 * (size_t)i + 1),
                      __ldg(a.xyz + 3 * (size_t)i + 2));
      q = __ldg(reinterpret_cast<const float4*>(a.rotation_raw) + i);
      if (j >= 0) {
        if (a.d_xyz) {
          m.x += __ldg(a.d_xyz + 3 * (size_t)j);
          m.y += __ldg(a.d_xyz + 3 * (size_t)j + 1);
          m.z += __ldg(a.d_xyz + 3 * (size_t)j + 2);
        }
        if (a.d_rot) {
          const float4 d = __ldg(reinterpret_cast<const float4*>(a.d_rot) + j);
          q.x += d.x; q.y += d.y; q.z += d.z; q.w += d.w;
        }
      }
      // torch.sigmoid / torch.exp / F.normalize(p=2, dim=1, eps=1e-12): v / max(||v||, eps)
      o = 1.0f / (1.0f + expf(-__ldg(a.opacity_raw + i)));
      if (a.isotropic) {
        const float e = expf(__ldg(a.scaling_raw + i));
        s = make_float3(e, e, e);
      } else {
        s = make_float3(expf(__ldg(a.scaling_raw + 3 * (size_t)i)), expf(__ldg(a.scaling_raw + 3 * (size_t)i + 1)),
                        expf(__ldg(a.scaling_raw + 3 * (size_t)i + 2)));
      }
      const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
      q.x /= n; q.y /= n; q.z /= n; q.w /= n;
    }
    a.means3D[3 * (size_t)i] = m.x; a.means3D[3 * (size_t)i + 1] = m.y; a.means3D[3 * (size_t)i + 2] = m.z;
    a.opacities[i] = o;
    a.scales[3 * (size_t)i] = s.x; a.scales[3 * (size_t)i + 1] = s.y; a.scales[3 * (size_t)i + 2] = s.z;
    reinterpret_cast<float4*>(a.rotations)[i] = q;
  }
  s_dyn[threadIdx.x] = code;
  __syncthreads();

  const int rows = min(ASM_BLOCK, a.P - first);
  // colour SH: out[r][k][c], k = 0 from f_dc, k >= 1 from f_rest
  {
    const int row = 3 * M, total = rows * row;
    float* out = a.shs + (size_t)first * row;
    for (int e = (int)threadIdx.x; e < total; e += ASM_BLOCK) {
      const int r = e / row, c = e - r * row;
      const int cd = s_dyn[r];
      float v = 0.f;
      if (cd != -2) {
        const size_t g = (size_t)(first + r);
        v = c < 3 ? __ldg(a.f_dc_color + g * 3 + c) : __ldg(a.f_rest_color + g * (size_t)(row - 3) + (c - 3));
        if (cd >= 0 && a.d_sh) v += __ldg(a.d_sh + (size_t)cd * row + c);
      }
      out[e] = v;
    }
  }
  // phase / amplitude SH: out[r][k][0] = phase, out[r][k][1] = amplitude
  {
    const int row = 2 * M, total = rows * row;
    float* out = a.shs_p + (size_t)first * row;
    for (int e = (int)threadIdx.x; e < total; e += ASM_BLOCK) {
      const int r = e / row, c = e - r * row;
      const int cd = s_dyn[r];
      float v = 0.f;
      if (cd != -2) {
        const size_t g = (size_t)(first + r);
        const int k = c >> 1;
        if (c & 1) v = k == 0 ? __ldg(a.f_dc_amp + g) : __ldg(a.f_rest_amp + g * (size_t)(M - 1) + (k - 1));
        else v = k == 0 ? __ldg(a.f_dc_phase + g) : __ldg(a.f_rest_phase + g * (size_t)(M - 1) + (k - 1));
        if (cd >= 0 && a.d_sh_p) v += __ldg(a.d_sh_p + (size_t)cd * row + c);
      }
      out[e] = v;
    }
  }
}

template <int MC>
__global__ void __launch_bounds__(ASM_BLOCK) assemble_bwd_kernel(GftAssembleArgs a, GftAssembleGrads g) {
  const int first = blockIdx.x * ASM_BLOCK;
  const int i = first + (int)threadIdx.x;
  const int M = MC ? MC : a.M;
  __shared__ int s_dyn[ASM_BLOCK];
  int code = -2;
  if (i < a.P) {
    const int j = a.dyn_index ? __ldg(a.dyn_index + i) : -1;
    const bool on = j >= 0 ? (a.include_dynamic != 0) : (a.include_static != 0);
    code = on ? (j >= 0 ? j : -1) : -2;
    float3 gm = make_float3(0.f, 0.f, 0.f), gs = gm;
    float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
    float go = 0.f;
    if (on) {
      gm = make_float3(__ldg(g.g_means3D + 3 * (size_t)i), __ldg(g.g_means3D + 3 * (size_t)i + 1),
                       __ldg(g.g_means3D + 3 * (size_t)i + 2));
      // sigmoid' = o (1 - o); exp' = exp
      const float o = __ldg(a.opacities + i);
      go = __ldg(g.g_opacities + i) * ((1.0f - o) * o);
      gs = make_float3(__ldg(g.g_scales + 3 * (size_t)i) * __ldg(a.scales + 3 * (size_t)i),
                       __ldg(g.g_scales + 3 * (size_t)i + 1) * __ldg(a.scales + 3 * (size_t)i + 1),
                       __ldg(g.g_scales + 3 * (size_t)i + 2) * __ldg(a.scales + 3 * (size_t)i + 2));
      // y = v / n, n = max(||v||, eps):  dv = (gy - y (y . gy)) / n   (dv = gy / eps below eps)
      float4 v = __ldg(reinterpret_cast<const float4*>(a.rotation_raw) + i);
      if (j >= 0 && a.d_rot) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(a.d_rot) + j);
        v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
      }
      const float4 gy = __ldg(reinterpret_cast<const float4*>(g.g_rotations) + i);
      const float nrm = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
      if (nrm > 1e-12f) {
        const float4 y = __ldg(reinterpret_cast<const float4*>(a.rotations) + i);
        const float dt = y.x * gy.x + y.y * gy.y + y.z * gy.z + y.w * gy.w;
        gq = make_float4((gy.x - y.x * dt) / nrm, (gy.y - y.y * dt) / nrm, (gy.z - y.z * dt) / nrm,
                         (gy.w - y.w * dt) / nrm);
      } else {
        gq = make_float4(gy.x / 1e-12f, gy.y / 1e-12f, gy.z / 1e-12f, gy.w / 1e-12f);
      }
    }
    g.g_xyz[3 * (size_t)i] = gm.x; g.g_xyz[3 * (size_t)i + 1] = gm.y; g.g_xyz[3 * (size_t)i + 2] = gm.z;
    g.g_opacity_raw[i] = go;
    if (a.isotropic) {
      g.g_scaling_raw[i] = gs.x + gs.y + gs.z;     // backward of repeat(1, 3)
    } else {
      g.g_scaling_raw[3 * (size_t)i] = gs.x; g.g_scaling_raw[3 * (size_t)i + 1] = gs.y;
      g.g_scaling_raw[3 * (size_t)i + 2] = gs.z;
    }
    reinterpret_cast<float4*>(g.g_rotation_raw)[i] = gq;
    if (code >= 0) {
      if (g.g_d_xyz) {
        g.g_d_xyz[3 * (size_t)code] = gm.x; g.g_d_xyz[3 * (size_t)code + 1] = gm.y;
        g.g_d_xyz[3 * (size_t)code + 2] = gm.z;
      }
      if (g.g_d_rot) reinterpret_cast<float4*>(g.g_d_rot)[code] = gq;
    } else if (j >= 0) {   // dynamic Gaussian of an excluded region: its deformation rows get zeros
      if (g.g_d_xyz) {
        g.g_d_xyz[3 * (size_t)j] = 0.f; g.g_d_xyz[3 * (size_t)j + 1] = 0.f; g.g_d_xyz[3 * (size_t)j + 2] = 0.f;
      }
      if (g.g_d_rot) reinterpret_cast<float4*>(g.g_d_rot)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      code = -3 - j;       // remembered for the SH rows below: excluded, deformation row j
    }
  }
  s_dyn[threadIdx.x] = code;
  __syncthreads();

  const int rows = min(ASM_BLOCK, a.P - first);
  {
    const int row = 3 * M, total = rows * row;
    const float* gin = g.g_shs + (size_t)first * row;
    for (int e = (int)threadIdx.x; e < total; e += ASM_BLOCK) {
      const int r = e / row, c = e - r * row;
      const int cd = s_dyn[r];
      const float v = cd >= -1 ? __ldg(gin + e) : 0.f;
      const size_t gi = (size_t)(first + r);
      if (c < 3) g.g_f_dc_color[gi * 3 + c] = v;
      else g.g_f_rest_color[gi * (size_t)(row - 3) + (c - 3)] = v;
      if (g.g_d_sh) {
        if (cd >= 0) g.g_d_sh[(size_t)cd * row + c] = v;
        else if (cd <= -3) g.g_d_sh[(size_t)(-3 - cd) * row + c] = 0.f;
      }
    }
  }
  {
    const int row = 2 * M, total = rows * row;
    const float* gin = g.g_shs_p + (size_t)first * row;
    for (int e = (int)threadIdx.x; e < total; e += ASM_BLOCK) {
      const int r = e / row, c = e - r * row;
      const int cd = s_dyn[r];
      const float v = cd >= -1 ? __ldg(gin + e) : 0.f;
      const size_t gi = (size_t)(first + r);
      const int k = c >> 1;
      if (c & 1) {
        if (k == 0) g.g_f_dc_amp[gi] = v; else g.g_f_rest_amp[gi * (size_t)(M - 1) + (k - 1)] = v;
      } else {
        if (k == 0) g.g_f_dc_phase[gi] = v; else g.g_f_rest_phase[gi * (size_t)(M - 1) + (k - 1)] = v;
      }
      if (g.g_d_sh_p) {
        if (cd >= 0) g.g_d_sh_p[(size_t)cd * row + c] = v;
        else if (cd <= -3) g.g_d_sh_p[(size_t)(-3 - cd) * row + c] = 0.f;
      }
    }
  }
}

// ---- Adam ------------------------------------------------------------------------------------
struct AdamScalars {
  float beta1, beta2, one_m_beta1, one_m_beta2, bc2_sqrt, eps;
  float neg_step[GFT_ADAM_MAX_SEGMENTS];   // -(lr / (1 - b1^t)) per segment, formed in double
};

// torch/optim/adam.py (_single_tensor_adam), in its operation order:
//   m <- m + (g - m) (1 - b1);  v <- v b2 + (1 - b2) g g;
//   denom = sqrt(v) / sqrt(1 - b2^t) + eps;  p <- p + (-lr / (1 - b1^t)) * (m / denom)
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float neg_step, const AdamScalars& k) {
  m = m + (g - m) * k.one_m_beta1;
  v = v * k.beta2 + k.one_m_beta2 * g * g;
  const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
  p = p + neg_step * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(GftAdamArgs a, AdamScalars k) {
  // every block sweeps every segment (= param group) with a grid stride; segment bounds are
  // multiples of 4, so a 128-bit access never straddles two learning rates
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int s = 0; s < a.n_segments; ++s) {
    const long long b4 = a.seg[s].begin >> 2, e4 = a.seg[s].end >> 2;
    const float neg_step = k.neg_step[s];
    for (long long q = b4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < e4; q += stride) {
      float4 p = reinterpret_cast<float4*>(a.param)[q];
      const float4 g = reinterpret_cast<const float4*>(a.grad)[q];
      float4 m = reinterpret_cast<float4*>(a.exp_avg)[q];
      float4 v = reinterpret_cast<float4*>(a.exp_avg_sq)[q];
      adam1(p.x, g.x, m.x, v.x, neg_step, k);
      adam1(p.y, g.y, m.y, v.y, neg_step, k);
      adam1(p.z, g.z, m.z, v.z, neg_step, k);
      adam1(p.w, g.w, m.w, v.w, neg_step, k);
      reinterpret_cast<float4*>(a.param)[q] = p;
      reinterpret_cast<float4*>(a.exp_avg)[q] = m;
      reinterpret_cast<float4*>(a.exp_avg_sq)[q] = v;
      if (a.zero_grad) reinterpret_cast<float4*>(a.grad)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

}  // namespace
}  // namespace gft

extern "C" {

int gft_assemble_forward(const GftAssembleArgs* a, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return gft::set_error(-1, "gft_assemble_forward: null args");
  if (a->P < 0 || a->M < 1) return gft::set_error(-1, "gft_assemble_forward: bad P / M");
  if (a->P == 0) return 0;
  if (!a->xyz || !a->opacity_raw || !a->scaling_raw || !a->rotation_raw || !a->f_dc_color ||
      !a->f_dc_phase || !a->f_dc_amp || (a->M > 1 && (!a->f_rest_color || !a->f_rest_phase || !a->f_rest_amp)))
    return gft::set_error(-1, "gft_assemble_forward: required parameter pointer is null");
  if (!a->means3D || !a->opacities || !a->scales || !a->rotations || !a->shs || !a->shs_p)
    return gft::set_error(-1, "gft_assemble_forward: required output pointer is null");
  const int blocks = (a->P + gft::ASM_BLOCK - 1) / gft::ASM_BLOCK;
  if (a->M == 16) gft::assemble_fwd_kernel<16><<<blocks, gft::ASM_BLOCK, 0, stream>>>(*a);
  else gft::assemble_fwd_kernel<0><<<blocks, gft::ASM_BLOCK, 0, stream>>>(*a);
  gft::note_launches(1);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

int gft_assemble_backward(const GftAssembleArgs* a, const GftAssembleGrads* g, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !g) return gft::set_error(-1, "gft_assemble_backward: null args");
  if (a->P < 0 || a->M < 1) return gft::set_error(-1, "gft_assemble_backward: bad P / M");
  if (a->P == 0) return 0;
  if (!a->rotation_raw || !a->opacities || !a->scales || !a->rotations)
    return gft::set_error(-1, "gft_assemble_backward: forward inputs / outputs missing");
  if (!g->g_means3D || !g->g_opacities || !g->g_scales || !g->g_rotations || !g->g_shs || !g->g_shs_p)
    return gft::set_error(-1, "gft_assemble_backward: incoming gradient pointer is null");
  if (!g->g_xyz || !g->g_opacity_raw || !g->g_scaling_raw || !g->g_rotation_raw || !g->g_f_dc_color ||
      !g->g_f_dc_phase || !g->g_f_dc_amp ||
      (a->M > 1 && (!g->g_f_rest_color || !g->g_f_rest_phase || !g->g_f_rest_amp)))
    return gft::set_error(-1, "gft_assemble_backward: outgoing gradient pointer is null");
  const int blocks = (a->P + gft::ASM_BLOCK - 1) / gft::ASM_BLOCK;
  if (a->M == 16) gft::assemble_bwd_kernel<16><<<blocks, gft::ASM_BLOCK, 0, stream>>>(*a, *g);
  else gft::assemble_bwd_kernel<0><<<blocks, gft::ASM_BLOCK, 0, stream>>>(*a, *g);
  gft::note_launches(1);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

int gft_adam_step(const GftAdamArgs* a, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return gft::set_error(-1, "gft_adam_step: null args");
  if (!a->param || !a->grad || !a->exp_avg || !a->exp_avg_sq)
    return gft::set_error(-1, "gft_adam_step: null buffer");
  if (a->n_segments < 0 || a->n_segments > GFT_ADAM_MAX_SEGMENTS)
    return gft::set_error(-1, "gft_adam_step: bad segment count");
  if (a->step < 1) return gft::set_error(-1, "gft_adam_step: step counts from 1");
  long long longest = 0;
  for (int s = 0; s < a->n_segments; ++s) {
    const GftAdamSegment& g = a->seg[s];
    if (g.begin < 0 || g.end < g.begin || (g.begin & 3) || (g.end & 3))
      return gft::set_error(-1, "gft_adam_step: segment bounds must be ordered multiples of 4");
    if (g.end - g.begin > longest) longest = g.end - g.begin;
  }
  if (longest == 0) return 0;
  // scalars in double, as torch computes them in Python floats, then rounded once
  const double bc1 = 1.0 - std::pow((double)a->beta1, (double)a->step);
  const double bc2 = 1.0 - std::pow((double)a->beta2, (double)a->step);
  gft::AdamScalars k;
  k.beta1 = a->beta1; k.beta2 = a->beta2;
  k.one_m_beta1 = (float)(1.0 - (double)a->beta1);
  k.one_m_beta2 = (float)(1.0 - (double)a->beta2);
  k.bc2_sqrt = (float)std::sqrt(bc2);
  k.eps = a->eps;
  long long blocks = (longest / 4 + 255) / 256;
  if (blocks > gft::sm_count() * 16) blocks = gft::sm_count() * 16;
  if (blocks < 1) blocks = 1;
  for (int s = 0; s < a->n_segments; ++s) k.neg_step[s] = (float)(-(a->seg[s].lr / bc1));
  gft::adam_kernel<<<(int)blocks, 256, 0, stream>>>(*a, k);
  gft::note_launches(1);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
