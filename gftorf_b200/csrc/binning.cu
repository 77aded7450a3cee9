// binning.cu — key duplication, tile|depth sort and tile-range identification.
//
// Replaces duplicateWithKeys (cuda_rasterizer/rasterizer_impl.cu:72-113), the
// cub::DeviceRadixSort::SortPairs call (rasterizer_impl.cu:334-339) and identifyTileRanges
// (rasterizer_impl.cu:118-140).  All of it is integer work and is bit-exact with the reference:
//   key   = (tile_id << 32) | float_bits(view_z)       value = Gaussian index
// emitted per Gaussian in (y outer, x inner) order at offset point_offsets[idx-1], then sorted
// stably on bits [0, 32 + ceil_log2-ish(T)) so equal keys keep ascending Gaussian index.
#include "common.cuh"
#include "kernels.h"
#include "radix_sort.cuh"

#include <cstdlib>
#include <cstring>

namespace gft {

// One block = 256 consecutive Gaussians; the block's instances are spread evenly over its
// threads (a Gaussian covering thousands of tiles is emitted by all 256 threads, with coalesced
// stores), instead of one thread looping over all tiles of its Gaussian.
__global__ void __launch_bounds__(GFT_BLOCK)
duplicate_keys_kernel(int P, const uint16_t* __restrict__ rect, const float* __restrict__ depths,
                      const uint32_t* __restrict__ point_offsets, uint64_t* __restrict__ keys,
                      uint32_t* __restrict__ values, int grid_x, uint32_t capacity, KeyFormat kf) {
  __shared__ uint32_t s_end[GFT_BLOCK];
  const int first = blockIdx.x * GFT_BLOCK;
  const int idx = first + threadIdx.x;
  const uint32_t base = first == 0 ? 0u : __ldg(point_offsets + first - 1);
  const int last = min(P, first + GFT_BLOCK) - 1;
  s_end[threadIdx.x] = __ldg(point_offsets + min(idx, last)) - base;
  __syncthreads();
  const uint32_t total = s_end[GFT_BLOCK - 1];
  for (uint32_t i = threadIdx.x; i < total; i += GFT_BLOCK) {
    // smallest t with s_end[t] > i
    int lo = 0, hi = GFT_BLOCK - 1;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int mid = (lo + hi) >> 1;
      if (s_end[mid] > i) hi = mid; else lo = mid + 1;
    }
    const int t = lo;
    const uint32_t k = i - (t > 0 ? s_end[t - 1] : 0u);
    const int g = first + t;
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(rect) + g);
    const uint32_t x0 = r.x & 0xffffu, y0 = r.x >> 16, x1 = r.y & 0xffffu;
    const uint32_t w = x1 - x0;
    const uint32_t ty = y0 + k / w, tx = x0 + k % w;
    uint64_t key = (uint64_t)(ty * (uint32_t)grid_x + tx);
    key <<= kf.depth_bits;
    // depth relative to the near plane, see KeyFormat (clamped: a NaN depth passes the frustum
    // test like in the reference and must not spill into the tile bits)
    const uint32_t zb = __float_as_uint(__ldg(depths + g));
    const uint32_t zmax = kf.depth_bits >= 32 ? 0xffffffffu : ((1u << kf.depth_bits) - 1u);
    key |= (uint64_t)(zb >= kf.depth_base ? min(zb - kf.depth_base, zmax) : 0u);
    if (base + i < capacity) {   // only ever false when a caller's size hint was too small
      keys[base + i] = key;
      values[base + i] = (uint32_t)g;
    }
  }
}

__global__ void identify_ranges_kernel(int R_cap, const uint32_t* __restrict__ d_R,
                                       const uint64_t* __restrict__ keys,
                                       uint2* __restrict__ ranges, int depth_bits) {
  const int R = d_R ? (int)min(__ldg(d_R), (uint32_t)R_cap) : R_cap;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R) return;
  const uint32_t cur = (uint32_t)(keys[idx] >> depth_bits);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[idx - 1] >> depth_bits);
    if (cur != prev) {
      ranges[prev].y = idx;
      ranges[cur].x = idx;
    }
  }
  if (idx == R - 1) ranges[cur].y = R;
}

void launch_duplicate_keys(int P, const int* /*radii*/, const uint16_t* rect, const float* depths,
                           const uint32_t* point_offsets, uint64_t* keys, uint32_t* values,
                           int grid_x, uint32_t capacity, KeyFormat kf, cudaStream_t stream) {
  const int blocks = (P + GFT_BLOCK - 1) / GFT_BLOCK;
  duplicate_keys_kernel<<<blocks, GFT_BLOCK, 0, stream>>>(P, rect, depths, point_offsets, keys,
                                                          values, grid_x, capacity, kf);
  note_launches(1);
}

void launch_identify_ranges(int R, const uint32_t* d_R, const uint64_t* keys, uint2* ranges,
                            KeyFormat kf, cudaStream_t stream) {
  if (R <= 0) return;
  identify_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>(R, d_R, keys, ranges, kf.depth_bits);
  note_launches(1);
}

KeyFormat key_format(float near_n, float far_n) {
  KeyFormat kf;
  kf.depth_bits = 32;
  kf.depth_base = 0u;
  static const bool full = [] {   // GFT_FULL_KEYS=1: the reference's 32-bit depth field (A/B runs)
    const char* e = std::getenv("GFT_FULL_KEYS");
    return e && e[0] == '1';
  }();
  if (!full && near_n > 0.f && far_n >= near_n && far_n < 3.0e38f) {
    uint32_t lo, hi;
    std::memcpy(&lo, &near_n, 4);
    std::memcpy(&hi, &far_n, 4);
    const uint32_t span = hi - lo;
    int b = 0;
    while (b < 32 && (span >> b) != 0u) ++b;
    kf.depth_bits = b == 0 ? 1 : b;
    kf.depth_base = lo;
  }
  return kf;
}

bool sort_result_in_out(int end_bit) { return sort_lands_in_out(end_bit); }

size_t sort_pairs_temp_bytes(int R) { return radix_sort_temp_bytes(R); }

int sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
               const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
               cudaStream_t stream, const uint32_t* d_R) {
  return radix_sort_pairs(d_temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, R, end_bit,
                          stream, d_R);
}

}  // namespace gft
