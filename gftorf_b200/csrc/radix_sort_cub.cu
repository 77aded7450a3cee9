// radix_sort_cub.cu — the library cross-check back-end: the same cub::DeviceRadixSort::SortPairs
// call the reference makes (cuda_rasterizer/rasterizer_impl.cu:334-339).  Not the default path;
// selected with GFT_SORT=cub to validate the hand-written sort and to time it against CUB's
// sm_100 tuning on the same box.
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>
#include <cstring>
#include "radix_sort.cuh"

namespace gft {

int radix_sort_passes(int end_bit);

size_t cub_sort_temp_bytes(int R) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, R > 0 ? R : 1);
  return bytes;
}

int cub_sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                   const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
                   cudaStream_t stream) {
  if (R <= 0) return 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_in, keys_out, vals_in,
                                                  vals_out, R, 0, end_bit, stream);
  return e == cudaSuccess ? 1 : -1;
}

int sort_backend() {
  const char* e = std::getenv("GFT_SORT");
  return (e && std::strcmp(e, "cub") == 0) ? 1 : 0;
}

size_t radix_sort_temp_bytes(int R) {
  const size_t a = own_sort_temp_bytes(R), b = cub_sort_temp_bytes(R);
  return a > b ? a : b;
}

// Result lands in (keys_out, vals_out) iff sort_lands_in_out(end_bit); otherwise in the input
// buffers (even number of ping-pong passes of the hand-written sort).
bool sort_lands_in_out(int end_bit) {
  return sort_backend() == 1 || (radix_sort_passes(end_bit) & 1);
}

int radix_sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                     const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
                     cudaStream_t stream, const uint32_t* d_R) {
  if (sort_backend() == 1) {
    if (d_R) return -3;  // the library call needs the count on the host
    return cub_sort_pairs(d_temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, R, end_bit,
                          stream);
  }
  return own_sort_pairs(d_temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, R, end_bit,
                        stream, d_R);
}

}  // namespace gft
