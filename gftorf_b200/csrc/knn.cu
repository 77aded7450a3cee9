// knn.cu — distCUDA2: mean squared distance to the 3 nearest other points, for sm_100a.
//
// Replaces SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-221) and its kernels
// coord2Morton (:45-70), boxMinMax (:78-117), boxMeanDist (:147-183), plus the two
// cub::DeviceReduce calls and the cub radix sort (:193-213).
//
// The result of the reference is the exact 3-NN mean: its box pruning only ever skips boxes that
// cannot hold a closer point, so the output does not depend on the visiting order, only on the
// squared-distance formula.  That formula is kept in the reference binary's contracted form
// (SASS of boxMeanDist: FMUL(dy,dy) -> FFMA(dx,dx,.) -> FFMA(dz,dz,.)), so the output is
// bit-identical.
//
// What changes for B200:
//   - no host round trips: the scene AABB stays on the device (the reference does 2 blocking D2H
//     copies, 1 cudaMalloc and 6 thrust::device_vector allocations per call);
//   - one fused min+max reduction instead of two cub passes over the points;
//   - the Morton-sorted points are gathered once into a contiguous float4 array, so the inner
//     search loop reads coalesced 16-byte records instead of points[indices[i]];
//   - the search is warp-cooperative: the 32 lanes of a warp are 32 consecutive points on the
//     Morton curve, they vote on which boxes any of them needs and then all scan those boxes in
//     lock-step through a shared-memory copy (broadcast reads, no divergence).
#include <float.h>
#include "common.cuh"
#include "kernels.h"
#include "radix_sort.cuh"

namespace gft {

namespace {

constexpr int KNN_BOX = 1024;  // BOX_SIZE, simple_knn.cu:18
constexpr int MM_BLOCKS = 1024; // capacity of the partial min/max array; the grid is 4 blocks per SM

struct KnnWs {
  float* partial;      // [MM_BLOCKS][6]
  float* aabb;         // [6] minx miny minz maxx maxy maxz
  uint64_t* keys_a;    // [P]
  uint64_t* keys_b;    // [P]
  uint32_t* vals_a;    // [P]
  uint32_t* vals_b;    // [P]
  float4* sorted;      // [P]
  float* boxes;        // [nboxes][8]
  void* sort_temp;
  size_t sort_temp_bytes;
  size_t total;
};

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

KnnWs knn_layout(char* base, int P) {
  KnnWs w;
  const size_t n = (size_t)(P > 0 ? P : 1);
  const size_t nboxes = (n + KNN_BOX - 1) / KNN_BOX;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base + off; off += up256(bytes); return p; };
  w.partial = (float*)take(MM_BLOCKS * 6 * sizeof(float));
  w.aabb = (float*)take(8 * sizeof(float));
  w.keys_a = (uint64_t*)take(n * 8);
  w.keys_b = (uint64_t*)take(n * 8);
  w.vals_a = (uint32_t*)take(n * 4);
  w.vals_b = (uint32_t*)take(n * 4);
  w.sorted = (float4*)take(n * 16);
  w.boxes = (float*)take(nboxes * 8 * sizeof(float));
  w.sort_temp_bytes = radix_sort_temp_bytes((int)n);
  w.sort_temp = take(w.sort_temp_bytes);
  w.total = off;
  return w;
}

// Component-wise min/max with init {0,0,0} (simple_knn.cu:191-200: the AABB always contains the
// origin).  CUDA's min/max on floats are fminf/fmaxf.
__global__ void __launch_bounds__(256) knn_minmax_kernel(int P, const float* __restrict__ pts,
                                                         float* __restrict__ partial) {
  float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __ldg(pts + 3 * (size_t)i + c);
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
  }
  __shared__ float s[8][6];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s[warp][c] = mn[c]; s[warp][3 + c] = mx[c]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float r = s[0][threadIdx.x];
    for (int w = 1; w < 8; ++w)
      r = threadIdx.x < 3 ? fminf(r, s[w][threadIdx.x]) : fmaxf(r, s[w][threadIdx.x]);
    partial[blockIdx.x * 6 + threadIdx.x] = r;
  }
}

__global__ void knn_minmax_final_kernel(int nblocks, const float* __restrict__ partial,
                                        float* __restrict__ aabb) {
  const int c = threadIdx.x;
  if (c >= 6) return;
  float r = partial[c];
  for (int b = 1; b < nblocks; ++b) r = c < 3 ? fminf(r, partial[b * 6 + c]) : fmaxf(r, partial[b * 6 + c]);
  aabb[c] = r;
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {  // simple_knn.cu:45-52
  x = (x | (x << 16)) & 0x030000FF;
  x = (x | (x << 8)) & 0x0300F00F;
  x = (x | (x << 4)) & 0x030C30C3;
  x = (x | (x << 2)) & 0x09249249;
  return x;
}

__global__ void knn_morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ aabb,
                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float x = __ldg(pts + 3 * (size_t)idx + 0);
  const float y = __ldg(pts + 3 * (size_t)idx + 1);
  const float z = __ldg(pts + 3 * (size_t)idx + 2);
  const float mnx = aabb[0], mny = aabb[1], mnz = aabb[2];
  const float mxx = aabb[3], mxy = aabb[4], mxz = aabb[5];
  // simple_knn.cu:54-61; float -> uint32 conversion truncates (saturating in hardware)
  const uint32_t cx = prep_morton((uint32_t)(((x - mnx) / (mxx - mnx)) * 1023.0f));
  const uint32_t cy = prep_morton((uint32_t)(((y - mny) / (mxy - mny)) * 1023.0f));
  const uint32_t cz = prep_morton((uint32_t)(((z - mnz) / (mxz - mnz)) * 1023.0f));
  keys[idx] = (uint64_t)(cx | (cy << 1) | (cz << 2));
  vals[idx] = (uint32_t)idx;
}

// Gather points in Morton order and build one AABB per 1024 consecutive sorted points.
__global__ void __launch_bounds__(KNN_BOX) knn_gather_boxes_kernel(int P, const float* __restrict__ pts,
                                                                   const uint32_t* __restrict__ order,
                                                                   float4* __restrict__ sorted,
                                                                   float* __restrict__ boxes) {
  const int idx = blockIdx.x * KNN_BOX + threadIdx.x;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (idx < P) {
    const uint32_t o = __ldg(order + idx);
    const float x = __ldg(pts + 3 * (size_t)o + 0);
    const float y = __ldg(pts + 3 * (size_t)o + 1);
    const float z = __ldg(pts + 3 * (size_t)o + 2);
    sorted[idx] = make_float4(x, y, z, __uint_as_float(o));
    mn[0] = mx[0] = x; mn[1] = mx[1] = y; mn[2] = mx[2] = z;
  }
  __shared__ float s[KNN_BOX / 32][6];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s[warp][c] = mn[c]; s[warp][3 + c] = mx[c]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float r = s[0][threadIdx.x];
    for (int w = 1; w < KNN_BOX / 32; ++w)
      r = threadIdx.x < 3 ? fminf(r, s[w][threadIdx.x]) : fmaxf(r, s[w][threadIdx.x]);
    boxes[blockIdx.x * 8 + threadIdx.x] = r;
  }
}

// squared distance in the reference's contracted form (SASS of boxMeanDist / updateKBest)
__device__ __forceinline__ float sqdist(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ void kbest3(float d, float& b0, float& b1, float& b2) {
  // updateKBest<3>, simple_knn.cu:131-145
  if (b0 > d) { const float t = b0; b0 = d; d = t; }
  if (b1 > d) { const float t = b1; b1 = d; d = t; }
  if (b2 > d) { b2 = d; }
}

// distBoxPoint, simple_knn.cu:119-129
__device__ __forceinline__ float box_dist(const float* __restrict__ b, float x, float y, float z) {
  float dx = 0.f, dy = 0.f, dz = 0.f;
  if (x < b[0] || x > b[3]) dx = fminf(fabsf(x - b[0]), fabsf(x - b[3]));
  if (y < b[1] || y > b[4]) dy = fminf(fabsf(y - b[1]), fabsf(y - b[4]));
  if (z < b[2] || z > b[5]) dz = fminf(fabsf(z - b[2]), fabsf(z - b[5]));
  return sqdist(dx, dy, dz);
}

constexpr int KNN_THREADS = 256;
constexpr int KNN_STAGE = 256;  // candidate points staged per step (4 KB)

__global__ void __launch_bounds__(KNN_THREADS)
knn_search_kernel(int P, const float4* __restrict__ sorted, const float* __restrict__ boxes,
                  int nboxes, float* __restrict__ out) {
  __shared__ float4 s_pts[KNN_THREADS / 32][KNN_STAGE / 8];  // per-warp staging: 32 points
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int idx = blockIdx.x * KNN_THREADS + threadIdx.x;
  const bool live = idx < P;
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) me = __ldg(sorted + idx);

  float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
  if (live) {
    for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); ++i) {
      if (i == idx) continue;
      const float4 q = __ldg(sorted + i);
      kbest3(sqdist(q.x - me.x, q.y - me.y, q.z - me.z), b0, b1, b2);
    }
  }
  const float reject = b2;  // simple_knn.cu:165
  b0 = b1 = b2 = FLT_MAX;

  float4* stage = s_pts[warp];
  for (int b = 0; b < nboxes; ++b) {
    bool want = false;
    if (live) {
      const float d = box_dist(boxes + 8 * (size_t)b, me.x, me.y, me.z);
      want = !(d > reject || d > b2);  // simple_knn.cu:172
    }
    if (!__any_sync(0xffffffffu, want)) continue;
    const int lo = b * KNN_BOX, hi = min(P, lo + KNN_BOX);
    for (int c = lo; c < hi; c += 32) {
      __syncwarp();
      if (c + (int)lane < hi) stage[lane] = __ldg(sorted + c + lane);
      __syncwarp();
      if (want) {
        const int cnt = min(32, hi - c);
        for (int j = 0; j < cnt; ++j) {
          if (c + j == idx) continue;
          const float4 q = stage[j];
          kbest3(sqdist(q.x - me.x, q.y - me.y, q.z - me.z), b0, b1, b2);
        }
      }
    }
  }
  if (live) out[__float_as_uint(me.w)] = __fdiv_rn(__fadd_rn(__fadd_rn(b0, b1), b2), 3.0f);
}

}  // namespace

size_t knn_workspace_bytes(int P) { return knn_layout(nullptr, P).total; }

int knn_dist2(const float* points, int P, float* out, char* workspace, cudaStream_t stream) {
  if (P <= 0) return 0;
  KnnWs w = knn_layout(workspace, P);
  const int mmb = min(min(MM_BLOCKS, 4 * sm_count()), (P + 255) / 256);
  knn_minmax_kernel<<<mmb, 256, 0, stream>>>(P, points, w.partial);
  knn_minmax_final_kernel<<<1, 32, 0, stream>>>(mmb, w.partial, w.aabb);
  knn_morton_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, w.aabb, w.keys_a, w.vals_a);
  const int rc = radix_sort_pairs(w.sort_temp, w.sort_temp_bytes, w.keys_a, w.keys_b, w.vals_a,
                                  w.vals_b, P, 30, stream);
  if (rc < 0) return rc;
  const uint32_t* order = sort_lands_in_out(30) ? w.vals_b : w.vals_a;
  const int nboxes = (P + KNN_BOX - 1) / KNN_BOX;
  knn_gather_boxes_kernel<<<nboxes, KNN_BOX, 0, stream>>>(P, points, order, w.sorted, w.boxes);
  knn_search_kernel<<<(P + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, stream>>>(
      P, w.sorted, w.boxes, nboxes, out);
  note_launches(5);
  return 0;
}

}  // namespace gft
