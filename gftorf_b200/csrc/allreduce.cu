// allreduce.cu — sum-allreduce of the flat gradient bucket through the NVSwitch (NVLS), for
// sm_100a.  C ABI: include/gftorf_train.h.
//
// The one exchange step of a data-parallel iteration: every rank holds the
// [P, 91] fp32 gradient bucket in SYMMETRIC memory (same allocation on every GPU, mapped through
// an NVLink multicast object; the host side obtains it from torch.distributed._symmetric_memory).
// Rank r owns the r-th slice of the bucket: it reads the slice through the multicast address with
// `multimem.ld_reduce` — the switch fetches the element from every GPU and returns the SUM, so the
// link carries one reduced copy — and writes the result back with `multimem.st`, which the switch
// replicates into every GPU's buffer.  Per GPU and direction the links carry ~S bytes in total,
// the lower bound for an allreduce; no staging buffers, no ring steps.  The caller brackets the
// launch with two cross-GPU barriers (data complete before, results visible after).
#include <cuda_runtime.h>
#include <cstdlib>

#include "../../include/gftorf_train.h"
#include "kernels.h"

namespace gft {
namespace {

__device__ __forceinline__ float4 mc_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

template <int AR_UNROLL>
__global__ void __launch_bounds__(256) nvls_allreduce_kernel(float4* __restrict__ mc, long long begin, long long end) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // AR_UNROLL independent reductions in flight per thread
  for (; i + (AR_UNROLL - 1) * stride < end; i += AR_UNROLL * stride) {
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) v[u] = mc_ld_reduce_add(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) mc_st(mc + i + u * stride, v[u]);
  }
  for (; i < end; i += stride) mc_st(mc + i, mc_ld_reduce_add(mc + i));
}

// ---- one kernel: cross-GPU barrier, reduce + broadcast of this rank's slice, cross-GPU barrier ----
// The barriers run over the symmetric-memory signal pads (one 32-bit word per (block, peer) pair):
// block b of rank r flips word [b][r] in every peer's pad from 0 to 1 (release, system scope) and
// then flips the words [b][p] of its own pad back from 1 to 0 (acquire), i.e. it waits for block b
// of every peer.  Barrier 1: a peer's block can only run when that peer's stream has finished
// every earlier kernel, so its gradients are complete.  Barrier 2: block b of every rank has
// stored its part; a rank's kernel ends when all its blocks have passed barrier 2, and every part
// of the bucket belongs to some block index, so the whole bucket is complete when the kernel ends.
// No host-launched barrier kernels around the exchange (they cost 2 x ~20 us of the 0.33 ms).
__device__ __forceinline__ void put_signal(uint32_t* addr) {
  uint32_t old;
  long long spins = 0;
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    if (++spins > (1ll << 31)) __trap();     // a lost peer must not hang the GPU
  } while (old != 0u);
}
__device__ __forceinline__ void wait_signal(uint32_t* addr) {
  uint32_t old;
  long long spins = 0;
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    if (++spins > (1ll << 31)) __trap();
  } while (old != 1u);
}
__device__ __forceinline__ void peer_barrier(uint32_t* const* pads, uint32_t word0, int rank, int world) {
  if ((int)threadIdx.x < world) {
    const int peer = (int)threadIdx.x;
    put_signal(pads[peer] + word0 + (size_t)blockIdx.x * world + rank);
    wait_signal(pads[rank] + word0 + (size_t)blockIdx.x * world + peer);
  }
}

template <int AR_UNROLL>
__global__ void __launch_bounds__(512)
nvls_allreduce_fused_kernel(float4* __restrict__ mc, long long begin, long long end,
                            uint32_t* const* __restrict__ pads, uint32_t word0, int rank, int world) {
  peer_barrier(pads, word0, rank, world);
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (AR_UNROLL - 1) * stride < end; i += AR_UNROLL * stride) {
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) v[u] = mc_ld_reduce_add(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) mc_st(mc + i + u * stride, v[u]);
  }
  for (; i < end; i += stride) mc_st(mc + i, mc_ld_reduce_add(mc + i));
  __threadfence_system();      // this thread's multicast stores are performed everywhere ...
  __syncthreads();             // ... for every thread of the block, before the block signals
  peer_barrier(pads, word0, rank, world);
}

// ---- the same exchange with plain peer-to-peer accesses (no multicast) ---------------------------
// Rank r sums slice r: its own values plus one 16-byte load from every peer's symmetric buffer,
// and stores the sum into every buffer.  Per GPU and direction the links carry 2 (N-1)/N S bytes
// against (1 + 1/N) S through the switch-side reduction: fewer at N = 2 (S against 1.5 S), more
// from N = 4 on — the autotuner decides.  Every element is summed by exactly one rank (own value
// first, then the peers in rank order) and broadcast, so all ranks end with identical bits.
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, float4 v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w) : "memory");
}

constexpr int P2P_MAX_WORLD = 8;
struct PeerBufs { float4* buf[P2P_MAX_WORLD]; };

template <int AR_UNROLL>
__global__ void __launch_bounds__(512)
p2p_allreduce_fused_kernel(const __grid_constant__ PeerBufs bufs, long long begin, long long end,
                           uint32_t* const* __restrict__ pads, uint32_t word0, int rank, int world) {
  peer_barrier(pads, word0, rank, world);
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < end; i += AR_UNROLL * stride) {
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < end) v[u] = ld_peer(bufs.buf[rank] + j);
    }
    for (int p = 0; p < world; ++p) {
      if (p == rank) continue;
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u) {
        const long long j = i + u * stride;
        if (j < end) {
          const float4 w = ld_peer(bufs.buf[p] + j);
          v[u].x += w.x; v[u].y += w.y; v[u].z += w.z; v[u].w += w.w;
        }
      }
    }
    for (int p = 0; p < world; ++p) {
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u) {
        const long long j = i + u * stride;
        if (j < end) st_peer(bufs.buf[p] + j, v[u]);
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  peer_barrier(pads, word0, rank, world);
}

// ---- the peer-to-peer exchange with posted stores only ------------------------------------------
// Remote LOADS are the slow half of the kernel above (530 GB/s per direction at N = 2 where the
// links carry 900).  Here nothing is loaded over the links: (1) every rank pushes the slices it
// does not own into region [rank] of the owner's symmetric scratch buffer, (2) barrier, (3) the
// owner sums its own values and the N-1 regions of its scratch (local memory) and pushes the sum
// into every rank's bucket, (4) barrier.  Same bytes per direction, 2 (N-1)/N bucket sizes.
// Element j of a slice is handled by the same (block, thread) on every rank, so the per-block
// barriers order exactly the data a block consumes.
struct PushBufs { float4* buf[P2P_MAX_WORLD]; float4* scratch[P2P_MAX_WORLD]; };

__global__ void __launch_bounds__(512)
push_allreduce_fused_kernel(const __grid_constant__ PushBufs b, long long per, long long total,
                            uint32_t* const* __restrict__ pads, uint32_t word0, int rank, int world) {
  peer_barrier(pads, word0, rank, world);            // every bucket is final, every scratch is free
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    const long long base = per * p;
    const long long n_p = min(per, total - base);
    float4* dst = b.scratch[p] + per * rank;
    const float4* src = b.buf[rank] + base;
    long long j = t0;
    for (; j + stride < n_p; j += 2 * stride) {      // two independent 16-byte copies in flight
      const float4 v0 = ld_peer(src + j), v1 = ld_peer(src + j + stride);
      st_peer(dst + j, v0);
      st_peer(dst + j + stride, v1);
    }
    if (j < n_p) st_peer(dst + j, ld_peer(src + j));
  }
  __threadfence_system();
  __syncthreads();
  peer_barrier(pads, word0, rank, world);            // all pushes have landed
  __syncthreads();
  {
    const long long base = per * rank;
    const long long n_r = min(per, total - base);
    for (long long j = t0; j < n_r; j += stride) {
      float4 acc = ld_peer(b.buf[rank] + base + j);
      for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        const float4 w = ld_peer(b.scratch[rank] + per * q + j);
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
      }
      for (int p = 0; p < world; ++p) st_peer(b.buf[p] + base + j, acc);
    }
  }
  __threadfence_system();
  __syncthreads();
  peer_barrier(pads, word0, rank, world);            // every bucket holds every sum
}

}  // namespace
}  // namespace gft

extern "C" {

int gft_push_allreduce_fused(void* const* buffer_ptrs_host, long long byte_offset, void* const* scratch_ptrs_host,
                             long long n_floats, int rank, int world, void* const* signal_pads_dev, int pad_words,
                             int blocks, gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!buffer_ptrs_host || !scratch_ptrs_host || !signal_pads_dev)
    return gft::set_error(-1, "gft_push_allreduce_fused: null pointer");
  if (world <= 0 || world > gft::P2P_MAX_WORLD || rank < 0 || rank >= world)
    return gft::set_error(-1, "gft_push_allreduce_fused: bad rank / world (at most 8 ranks)");
  if (n_floats < 0 || (n_floats & 3) || (byte_offset & 15))
    return gft::set_error(-1, "gft_push_allreduce_fused: length must be a multiple of 4 floats, offset of 16 bytes");
  gft::PushBufs pb;
  for (int p = 0; p < gft::P2P_MAX_WORLD; ++p) {
    char* bb = p < world ? static_cast<char*>(buffer_ptrs_host[p]) : nullptr;
    char* sb = p < world ? static_cast<char*>(scratch_ptrs_host[p]) : nullptr;
    if (p < world && (!bb || !sb || (reinterpret_cast<uintptr_t>(bb) & 15) || (reinterpret_cast<uintptr_t>(sb) & 15)))
      return gft::set_error(-1, "gft_push_allreduce_fused: buffer pointers must be non-null and 16-byte aligned");
    pb.buf[p] = bb ? reinterpret_cast<float4*>(bb + byte_offset) : nullptr;
    pb.scratch[p] = reinterpret_cast<float4*>(sb);
  }
  const long long total = n_floats >> 2;
  const long long per = (total + world - 1) / world;
  const int word0 = pad_words / 2;
  int max_blocks = (pad_words - word0) / world;
  if (max_blocks < 1) return gft::set_error(-1, "gft_push_allreduce_fused: signal pad too small");
  long long want = blocks > 0 ? blocks : (long long)gft::sm_count();
  const long long need = (per + 512 - 1) / 512;
  if (want > need) want = need > 0 ? need : 1;
  if (want > max_blocks) want = max_blocks;
  gft::push_allreduce_fused_kernel<<<(int)want, 512, 0, stream>>>(
      pb, per, total, reinterpret_cast<uint32_t* const*>(signal_pads_dev), (uint32_t)word0, rank, world);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(err));
  gft::note_launches(1);
  return 0;
}

int gft_p2p_allreduce_fused(void* const* buffer_ptrs_host, long long byte_offset, long long n_floats, int rank,
                            int world, void* const* signal_pads_dev, int pad_words, int blocks, int unroll,
                            gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!buffer_ptrs_host || !signal_pads_dev) return gft::set_error(-1, "gft_p2p_allreduce_fused: null pointer");
  if (world <= 0 || world > gft::P2P_MAX_WORLD || rank < 0 || rank >= world)
    return gft::set_error(-1, "gft_p2p_allreduce_fused: bad rank / world (at most 8 ranks)");
  if (n_floats < 0 || (n_floats & 3) || (byte_offset & 15))
    return gft::set_error(-1, "gft_p2p_allreduce_fused: length must be a multiple of 4 floats, offset of 16 bytes");
  gft::PeerBufs pb;
  for (int p = 0; p < gft::P2P_MAX_WORLD; ++p) {
    char* base = p < world ? static_cast<char*>(buffer_ptrs_host[p]) : nullptr;
    if (p < world && (!base || (reinterpret_cast<uintptr_t>(base) & 15)))
      return gft::set_error(-1, "gft_p2p_allreduce_fused: peer buffer pointers must be non-null and 16-byte aligned");
    pb.buf[p] = base ? reinterpret_cast<float4*>(base + byte_offset) : nullptr;
  }
  const long long total = n_floats >> 2;
  const long long per = (total + world - 1) / world;
  const long long begin = per * rank, end = begin + per < total ? begin + per : total;
  const int word0 = pad_words / 2;
  int max_blocks = (pad_words - word0) / world;
  if (max_blocks < 1) return gft::set_error(-1, "gft_p2p_allreduce_fused: signal pad too small");
  // EVERY rank must launch the same number of blocks (block b pairs with block b of its peers)
  long long want = blocks > 0 ? blocks : (long long)gft::sm_count();
  const long long need = (per + 512 - 1) / 512;
  if (want > need) want = need > 0 ? need : 1;
  if (want > max_blocks) want = max_blocks;
  uint32_t* const* pads = reinterpret_cast<uint32_t* const*>(signal_pads_dev);
  const long long e = end > begin ? end : begin;
  if (unroll == 4) gft::p2p_allreduce_fused_kernel<4><<<(int)want, 512, 0, stream>>>(pb, begin, e, pads, (uint32_t)word0, rank, world);
  else if (unroll == 1) gft::p2p_allreduce_fused_kernel<1><<<(int)want, 512, 0, stream>>>(pb, begin, e, pads, (uint32_t)word0, rank, world);
  else gft::p2p_allreduce_fused_kernel<2><<<(int)want, 512, 0, stream>>>(pb, begin, e, pads, (uint32_t)word0, rank, world);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(err));
  gft::note_launches(1);
  return 0;
}

int gft_nvls_allreduce_fused(float* multicast_ptr, long long n_floats, int rank, int world,
                             void* const* signal_pads_dev, int pad_words, int blocks, int unroll,
                             gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!multicast_ptr || !signal_pads_dev) return gft::set_error(-1, "gft_nvls_allreduce_fused: null pointer");
  if (world <= 0 || world > 32 || rank < 0 || rank >= world) return gft::set_error(-1, "gft_nvls_allreduce_fused: bad rank / world");
  if (n_floats < 0 || (n_floats & 3) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15))
    return gft::set_error(-1, "gft_nvls_allreduce_fused: length must be a multiple of 4 floats, pointer 16-byte aligned");
  const long long total = n_floats >> 2;
  const long long per = (total + world - 1) / world;
  const long long begin = per * rank, end = begin + per < total ? begin + per : total;
  // the upper half of every signal pad is ours (the lower half serves the host-side barriers)
  const int word0 = pad_words / 2;
  int max_blocks = (pad_words - word0) / world;
  if (max_blocks < 1) return gft::set_error(-1, "gft_nvls_allreduce_fused: signal pad too small");
  // EVERY rank must launch the same number of blocks (block b pairs with block b of its peers),
  // so the count depends only on the arguments all ranks share
  long long want = blocks > 0 ? blocks : (long long)gft::sm_count() * 2;
  const long long need = (per + 512 - 1) / 512;
  if (want > need) want = need > 0 ? need : 1;
  if (want > max_blocks) want = max_blocks;
  float4* mc = reinterpret_cast<float4*>(multicast_ptr);
  uint32_t* const* pads = reinterpret_cast<uint32_t* const*>(signal_pads_dev);
  if (unroll == 8) gft::nvls_allreduce_fused_kernel<8><<<(int)want, 512, 0, stream>>>(mc, begin, end > begin ? end : begin, pads, (uint32_t)word0, rank, world);
  else if (unroll == 2) gft::nvls_allreduce_fused_kernel<2><<<(int)want, 512, 0, stream>>>(mc, begin, end > begin ? end : begin, pads, (uint32_t)word0, rank, world);
  else gft::nvls_allreduce_fused_kernel<4><<<(int)want, 512, 0, stream>>>(mc, begin, end > begin ? end : begin, pads, (uint32_t)word0, rank, world);
  gft::note_launches(1);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

int gft_nvls_allreduce_sum(float* multicast_ptr, long long n_floats, int rank, int world,
                           gft_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!multicast_ptr) return gft::set_error(-1, "gft_nvls_allreduce_sum: null multicast pointer");
  if (world <= 0 || rank < 0 || rank >= world) return gft::set_error(-1, "gft_nvls_allreduce_sum: bad rank / world");
  if (n_floats < 0 || (n_floats & 3) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15))
    return gft::set_error(-1, "gft_nvls_allreduce_sum: length must be a multiple of 4 floats, pointer 16-byte aligned");
  const long long total = n_floats >> 2;
  const long long per = (total + world - 1) / world;
  const long long begin = per * rank, end = begin + per < total ? begin + per : total;
  if (begin >= end) return 0;
  static const int env_blocks = [] { const char* e = std::getenv("GFT_NVLS_BLOCKS"); return e ? std::atoi(e) : 0; }();
  static const int env_unroll = [] { const char* e = std::getenv("GFT_NVLS_UNROLL"); return e ? std::atoi(e) : 0; }();
  const int unroll = env_unroll == 8 ? 8 : (env_unroll == 2 ? 2 : 4);
  long long blocks = (end - begin + 256 * unroll - 1) / (256 * unroll);
  const long long cap = env_blocks > 0 ? env_blocks : (long long)gft::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  float4* mc = reinterpret_cast<float4*>(multicast_ptr);
  if (unroll == 8) gft::nvls_allreduce_kernel<8><<<(int)blocks, 256, 0, stream>>>(mc, begin, end);
  else if (unroll == 2) gft::nvls_allreduce_kernel<2><<<(int)blocks, 256, 0, stream>>>(mc, begin, end);
  else gft::nvls_allreduce_kernel<4><<<(int)blocks, 256, 0, stream>>>(mc, begin, end);
  gft::note_launches(1);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gft::set_error(-2, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
