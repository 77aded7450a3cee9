// densify.cu — adaptive density control as a plan + one gather per parameter group, for sm_100a
// (SURVEY.md §8f row f2).  C ABI: include/gftorf_train.h.
//
// Replaces GaussianModel.densify_and_prune (scene/gaussian_model.py:631-646) and what it calls:
// densify_and_clone (:607-629), densify_and_split (:568-605), prune_points (:493-513) and the
// optimizer-state surgery (_prune_optimizer :473-491, cat_tensors_to_optimizer :515-536).  The
// reference rebuilds every one of its 11 parameter tensors and both Adam moments three times (two
// torch.cat rounds and two boolean-mask compactions: ~12 full-size reallocations and copies per
// tensor).  Here the outcome of the whole sequence is decided first, per ORIGINAL Gaussian:
//
//   clone   c = |g| >= max_grad  and  max_scale <= percent_dense * extent
//   split   s =  g  >= max_grad  and  max_scale >  percent_dense * extent          (c, s disjoint)
//   pruned  as original / clone : opacity < min_opacity  [or scale out of (0.001, 0.05) * extent]
//           as split child      : the same tests on the child's scale  exp(log(scale / 1.6))
//
// and the final order the reference ends up with is
//   [ originals with !s and !pruned | clones with !pruned | children copy 0 | children copy 1 ]
// (each part in ascending index order).  Three small kernels turn the flags into a plan
// (source index, kind, noise row per output Gaussian); one gather kernel per parameter group then
// writes the new parameter and both moments exactly once.
#include <cuda_runtime.h>
#include <cstring>

#include "../../include/gftorf_train.h"
#include "kernels.h"

namespace gft {
namespace {

constexpr int DB = 256;

struct Flags { bool keep, clone, split, child; };

__device__ __forceinline__ Flags classify(const GftDensifyPlanArgs& a, int i) {
  // grads = accum / denom with NaN -> 0 (gaussian_model.py:632-633)
  float g = __ldg(a.grad_accum + i) / __ldg(a.denom + i);
  if (g != g) g = 0.f;
  float ms;
  if (a.isotropic) {
    ms = expf(__ldg(a.scaling_raw + i));
  } else {
    ms = fmaxf(fmaxf(expf(__ldg(a.scaling_raw + 3 * (size_t)i)), expf(__ldg(a.scaling_raw + 3 * (size_t)i + 1))),
               expf(__ldg(a.scaling_raw + 3 * (size_t)i + 2)));
  }
  const float dense = a.percent_dense * a.extent;
  const bool c = (fabsf(g) >= a.max_grad) && (ms <= dense);
  const bool s = (g >= a.max_grad) && (ms > dense);
  const float op = 1.0f / (1.0f + expf(-__ldg(a.opacity_raw + i)));
  bool pr = op < a.min_opacity, prc = pr;
  if (a.size_prune) {
    pr = pr || ms > 0.05f * a.extent || ms < 0.001f * a.extent;
    // a child's scale goes through log(scale / (0.8 * 2)) and back through exp (:585, :125)
    float mc;
    if (a.isotropic) {
      mc = expf(logf(expf(__ldg(a.scaling_raw + i)) / 1.6f));
    } else {
      mc = fmaxf(fmaxf(expf(logf(expf(__ldg(a.scaling_raw + 3 * (size_t)i)) / 1.6f)),
                       expf(logf(expf(__ldg(a.scaling_raw + 3 * (size_t)i + 1)) / 1.6f))),
                 expf(logf(expf(__ldg(a.scaling_raw + 3 * (size_t)i + 2)) / 1.6f)));
    }
    prc = prc || mc > 0.05f * a.extent || mc < 0.001f * a.extent;
  }
  Flags f;
  f.keep = !s && !pr;
  f.clone = c && !pr;
  f.split = s;
  f.child = s && !prc;
  return f;
}

// The four 0/1 flags are scanned together: a block has 256 threads, so each running count fits in
// 9 bits and the four counters share one 64-bit word (16 bits each).
__device__ __forceinline__ unsigned long long pack(const Flags& f) {
  return (unsigned long long)f.keep | ((unsigned long long)f.clone << 16) |
         ((unsigned long long)f.split << 32) | ((unsigned long long)f.child << 48);
}

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* s_w,
                                                                   unsigned long long& total) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += n;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  unsigned long long base = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < DB / 32; ++w) {
    if ((unsigned)w < warp) base += s_w[w];
    total += s_w[w];
  }
  __syncthreads();
  return base + incl - v;
}

// K1: per-block counts of the four flags
__global__ void __launch_bounds__(DB) densify_count_kernel(GftDensifyPlanArgs a, uint32_t* __restrict__ block_counts) {
  __shared__ unsigned long long s_w[DB / 32];
  const int i = blockIdx.x * DB + threadIdx.x;
  Flags f = {false, false, false, false};
  if (i < a.P) f = classify(a, i);
  unsigned long long total;
  block_exclusive_scan(pack(f), s_w, total);
  if (threadIdx.x == 0) {
    uint32_t* o = block_counts + 4 * (size_t)blockIdx.x;
    o[0] = (uint32_t)(total & 0xffff); o[1] = (uint32_t)((total >> 16) & 0xffff);
    o[2] = (uint32_t)((total >> 32) & 0xffff); o[3] = (uint32_t)((total >> 48) & 0xffff);
  }
}

// K2: one block turns the per-block counts into exclusive offsets (in place) and the totals
__global__ void __launch_bounds__(1024) densify_offsets_kernel(uint32_t* __restrict__ block_counts, int nblocks,
                                                               int32_t* __restrict__ counts) {
  __shared__ uint32_t s_part[4][1024];
  const int t = threadIdx.x;
  const int per = (nblocks + 1023) / 1024;
  const int b0 = t * per, b1 = min(nblocks, b0 + per);
  uint32_t sum[4] = {0, 0, 0, 0};
  for (int b = b0; b < b1; ++b)
    for (int k = 0; k < 4; ++k) sum[k] += block_counts[4 * (size_t)b + k];
  for (int k = 0; k < 4; ++k) s_part[k][t] = sum[k];
  __syncthreads();
  // Hillis-Steele over the 1024 partials (tiny)
  for (int o = 1; o < 1024; o <<= 1) {
    uint32_t add[4];
    for (int k = 0; k < 4; ++k) add[k] = t >= o ? s_part[k][t - o] : 0u;
    __syncthreads();
    for (int k = 0; k < 4; ++k) s_part[k][t] += add[k];
    __syncthreads();
  }
  uint32_t run[4];
  for (int k = 0; k < 4; ++k) run[k] = s_part[k][t] - sum[k];
  for (int b = b0; b < b1; ++b)
    for (int k = 0; k < 4; ++k) {
      const uint32_t c = block_counts[4 * (size_t)b + k];
      block_counts[4 * (size_t)b + k] = run[k];
      run[k] += c;
    }
  if (t == 1023) {
    const int n_keep = (int)s_part[0][t], n_clone = (int)s_part[1][t], n_split = (int)s_part[2][t],
              n_child = (int)s_part[3][t];
    counts[0] = n_keep; counts[1] = n_clone; counts[2] = n_split; counts[3] = n_child;
    counts[4] = n_keep + n_clone + 2 * n_child;   // the new number of Gaussians
  }
}

// K3: the plan.  plan_src[k] = original index, plan_kind[k] = 0 keep / 1 clone / 2 child,
// plan_noise[k] = row of the normal samples for a child (copy * n_split + rank among the split).
__global__ void __launch_bounds__(DB) densify_plan_kernel(GftDensifyPlanArgs a, const uint32_t* __restrict__ block_offsets,
                                                          const int32_t* __restrict__ counts) {
  __shared__ unsigned long long s_w[DB / 32];
  const int i = blockIdx.x * DB + threadIdx.x;
  Flags f = {false, false, false, false};
  if (i < a.P) f = classify(a, i);
  unsigned long long total;
  const unsigned long long ex = block_exclusive_scan(pack(f), s_w, total);
  if (i >= a.P) return;
  const uint32_t* bo = block_offsets + 4 * (size_t)blockIdx.x;
  const int k_rank = (int)(bo[0] + (uint32_t)(ex & 0xffff));
  const int c_rank = (int)(bo[1] + (uint32_t)((ex >> 16) & 0xffff));
  const int s_rank = (int)(bo[2] + (uint32_t)((ex >> 32) & 0xffff));
  const int ch_rank = (int)(bo[3] + (uint32_t)((ex >> 48) & 0xffff));
  const int n_keep = counts[0], n_clone = counts[1], n_split = counts[2], n_child = counts[3];
  if (f.keep) {
    a.plan_src[k_rank] = i; a.plan_kind[k_rank] = 0; a.plan_noise[k_rank] = -1;
  }
  if (f.clone) {
    const int k = n_keep + c_rank;
    a.plan_src[k] = i; a.plan_kind[k] = 1; a.plan_noise[k] = -1;
  }
  if (f.child) {
    const int k0 = n_keep + n_clone + ch_rank, k1 = k0 + n_child;
    a.plan_src[k0] = i; a.plan_kind[k0] = 2; a.plan_noise[k0] = s_rank;
    a.plan_src[k1] = i; a.plan_kind[k1] = 2; a.plan_noise[k1] = n_split + s_rank;
  }
  if (f.split && a.split_src) a.split_src[s_rank] = i;   // the selected Gaussians, in index order
}

// K4: gather one parameter group and its two Adam moments into the new Gaussian set.
//   mode 0: plain rows;  mode 1: xyz (children: R(rotation) * noise + xyz);
//   mode 2: scaling (children: log(exp(s) / 1.6)).
__global__ void __launch_bounds__(DB) densify_gather_kernel(GftDensifyApplyArgs a) {
  const long long total = (long long)a.P_new * a.width;
  const long long e = (long long)blockIdx.x * DB + threadIdx.x;
  if (e >= total) return;
  const int k = (int)(e / a.width), c = (int)(e - (long long)k * a.width);
  const int src = __ldg(a.plan_src + k);
  const int kind = __ldg(a.plan_kind + k);
  const size_t si = (size_t)src * a.width + c;
  float v = __ldg(a.param_in + si);
  if (kind == 2) {
    if (a.mode == 1) {
      // samples = N(0,1) * scale; new_xyz = R(q / |q|) * samples + xyz   (:579-582)
      const int row = __ldg(a.plan_noise + k);
      const float n0 = __ldg(a.noise + 3 * (size_t)row), n1 = __ldg(a.noise + 3 * (size_t)row + 1),
                  n2 = __ldg(a.noise + 3 * (size_t)row + 2);
      const float4 q = __ldg(reinterpret_cast<const float4*>(a.rotation_raw) + src);
      const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      const float r = q.x / nrm, x = q.y / nrm, y = q.z / nrm, z = q.w / nrm;
      float R0, R1, R2;
      if (c == 0) { R0 = 1.f - 2.f * (y * y + z * z); R1 = 2.f * (x * y - r * z); R2 = 2.f * (x * z + r * y); }
      else if (c == 1) { R0 = 2.f * (x * y + r * z); R1 = 1.f - 2.f * (x * x + z * z); R2 = 2.f * (y * z - r * x); }
      else { R0 = 2.f * (x * z - r * y); R1 = 2.f * (y * z + r * x); R2 = 1.f - 2.f * (x * x + y * y); }
      v = (R0 * n0 + R1 * n1 + R2 * n2) + v;
    } else if (a.mode == 2) {
      v = logf(expf(v) / 1.6f);
    }
  }
  a.param_out[e] = v;
  const bool carry = kind == 0;   // new Gaussians start with zero moments (:527-528)
  if (a.exp_avg_out) a.exp_avg_out[e] = carry ? __ldg(a.exp_avg_in + si) : 0.f;
  if (a.exp_avg_sq_out) a.exp_avg_sq_out[e] = carry ? __ldg(a.exp_avg_sq_in + si) : 0.f;
}

}  // namespace
}  // namespace gft

extern "C" {

size_t gft_densify_workspace_bytes(int P) {
  const size_t blocks = ((size_t)(P > 0 ? P : 1) + gft::DB - 1) / gft::DB;
  return blocks * 4 * sizeof(uint32_t) + 256;
}

int gft_densify_plan(const GftDensifyPlanArgs* a, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return gft::set_error(-1, "gft_densify_plan: null args");
  if (a->P <= 0) return gft::set_error(-1, "gft_densify_plan: P must be positive");
  if (!a->grad_accum || !a->denom || !a->opacity_raw || !a->scaling_raw)
    return gft::set_error(-1, "gft_densify_plan: null input");
  if (!a->plan_src || !a->plan_kind || !a->plan_noise || !a->counts || !a->workspace)
    return gft::set_error(-1, "gft_densify_plan: null output / workspace");
  const int blocks = (a->P + gft::DB - 1) / gft::DB;
  if (blocks > 65535 * 16) return gft::set_error(-1, "gft_densify_plan: P too large");
  uint32_t* bc = reinterpret_cast<uint32_t*>(a->workspace);
  gft::densify_count_kernel<<<blocks, gft::DB, 0, stream>>>(*a, bc);
  gft::densify_offsets_kernel<<<1, 1024, 0, stream>>>(bc, blocks, a->counts);
  gft::densify_plan_kernel<<<blocks, gft::DB, 0, stream>>>(*a, bc, a->counts);
  gft::note_launches(3);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

int gft_densify_apply(const GftDensifyApplyArgs* a, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return gft::set_error(-1, "gft_densify_apply: null args");
  if (a->P_new < 0 || a->width <= 0) return gft::set_error(-1, "gft_densify_apply: bad P_new / width");
  if (a->P_new == 0) return 0;
  if (!a->plan_src || !a->plan_kind || !a->param_in || !a->param_out)
    return gft::set_error(-1, "gft_densify_apply: null pointer");
  if ((a->exp_avg_out && !a->exp_avg_in) || (a->exp_avg_sq_out && !a->exp_avg_sq_in))
    return gft::set_error(-1, "gft_densify_apply: moment output without input");
  // `noise` may be NULL when nothing was selected for splitting (no child reads it)
  if (a->mode == 1 && (a->width != 3 || !a->rotation_raw || !a->plan_noise))
    return gft::set_error(-1, "gft_densify_apply: xyz mode needs width 3, rotation_raw and plan_noise");
  if (a->mode < 0 || a->mode > 2) return gft::set_error(-1, "gft_densify_apply: mode must be 0..2");
  const long long total = (long long)a->P_new * a->width;
  const long long blocks = (total + gft::DB - 1) / gft::DB;
  gft::densify_gather_kernel<<<(unsigned)blocks, gft::DB, 0, stream>>>(*a);
  gft::note_launches(1);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
