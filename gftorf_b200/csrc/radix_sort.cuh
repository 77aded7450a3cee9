// radix_sort.cuh — stable LSD radix sort of (u64 key, u32 value) pairs on bits [0, end_bit).
//
// Two back-ends behind one interface (selected by gft::sort_backend(), env GFT_SORT):
//   "own"  hand-written 8-bit onesweep-style sort for sm_100a (radix_sort_own.cu)
//   "cub"  cub::DeviceRadixSort::SortPairs — the same library call the reference makes
//          (rasterizer_impl.cu:334-339); kept as the cross-check for the hand-written path.
// Both are stable, so both yield exactly the reference's permutation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace gft {

size_t radix_sort_temp_bytes(int R);
// d_R (optional): device pointer to the actual pair count, <= R; R is then only the capacity the
// grids and the temp storage are sized for (own back-end only).
int radix_sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                     const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
                     cudaStream_t stream, const uint32_t* d_R = nullptr);

// own back-end (radix_sort_own.cu)
size_t own_sort_temp_bytes(int R);
int own_sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                   const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
                   cudaStream_t stream, const uint32_t* d_R = nullptr);

// cub back-end (radix_sort_cub.cu)
size_t cub_sort_temp_bytes(int R);
int cub_sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                   const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
                   cudaStream_t stream);

// 0 = own, 1 = cub
int sort_backend();
int radix_sort_passes(int end_bit);
// true: the sorted pairs end up in (keys_out, vals_out); false: back in the input buffers
bool sort_lands_in_out(int end_bit);

}  // namespace gft
