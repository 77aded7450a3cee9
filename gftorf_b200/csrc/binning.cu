// binning.cu — tile-segmented binning: per-tile instance counts -> tile ranges -> scatter of the
// (depth, Gaussian) entries into their tile's segment -> one in-shared-memory sort per segment.
//
// Replaces duplicateWithKeys (cuda_rasterizer/rasterizer_impl.cu:72-113), the
// cub::DeviceRadixSort::SortPairs call over all R instances (rasterizer_impl.cu:334-339: 6 global
// passes of 24 B/instance + histogram) and identifyTileRanges + its memset
// (rasterizer_impl.cu:118-140,341).  All of it is integer work and bit-exact with the reference:
// the reference sorts keys (tile << 32) | float_bits(view_z) with a STABLE radix sort over
// instances emitted in ascending Gaussian order, so inside a tile the list is ordered by
// (float_bits(view_z), Gaussian index).  Here
//   1. the preprocess kernel counts the instances of every tile (tile_counts, one RED each).  A
//      tile has S sub-counters (Gaussian index mod S): ~1000 instances per tile hammering ONE
//      address serialise in the L2 (measured 0.165 ms for 2.4 M atomics on 2400 addresses, the
//      same whether 2 or 16 views are batched); S = 16 spreads them over 16 addresses,
//   2. tile_scan_kernel turns the counts into sub-bin starts, `ranges` (an exclusive scan over
//      tiles IS the reference's ranges array) and the instance count R,
//   3. scatter_entries_kernel writes every instance's 64-bit entry (depth bits << 32 | index)
//      into its tile's segment at a slot taken from the sub-bin's atomic cursor (arbitrary order),
//   4. tile_sort_kernel sorts each segment: a stable 4-pass LSD radix sort on the 32 depth bits
//      in shared memory, then a check for equal depths in the wrong index order (only then the
//      whole 64-bit entry is sorted, with the bitonic network that also serves segments too long
//      for shared memory).  The result is a total order on (depth bits, index), so the arbitrary
//      slot order of step 3 does not matter and the list is the reference's.
// HBM traffic: 8 B written + 8 B read + 12 B written per instance (28 B) instead of
// 12 + 8 + 6 x 24 + 8 = 172 B for emit + histogram + radix passes + ranges.
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace gft {

namespace {

typedef unsigned long long u64;

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;                       // one 16-byte load / store per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ranges[t] = (starts[t*S], starts[(t+1)*S]), with (0,0) for empty tiles — the reference's memset
// + identifyTileRanges never touch those (rasterizer_impl.cu:118-140,341) — and the launch order of
// the blend kernels.  Run by ONE block (the scan block that finishes last): T <= 16 views x 8160
// tiles.  `starts` was written by other blocks: read through the L2.
__device__ __forceinline__ void tile_ranges_and_order(const uint32_t* starts, int T, int S,
                                                      uint2* __restrict__ ranges, uint32_t* __restrict__ order,
                                                      uint32_t* s_bk) {
  const uint32_t tid = threadIdx.x;
  // Launch order of the blend kernels: longest tile lists first.  A blend block's time is
  // proportional to its tile's list length, the lists differ by 2-4x, and a grid is only ~5 blocks
  // per resident slot deep, so in index order a long tile that starts late leaves most SMs idle at
  // the end of the kernel.  128 buckets on a quarter-octave scale; the order inside a bucket is
  // whatever the atomics give (it only decides which SM runs a tile, never a result).
  auto bucket = [](uint32_t len) -> uint32_t {
    if (len < 4u) return len;
    const uint32_t e = 31u - __clz(len);
    return min(127u, e * 4u + ((len >> (e - 2u)) & 3u));
  };
  if (tid < 128) s_bk[tid] = 0u;
  __syncthreads();
  for (int t = tid; t < T; t += SCAN_THREADS) {
    const uint32_t x = __ldcg(starts + (size_t)t * S), y = __ldcg(starts + (size_t)(t + 1) * S);
    ranges[t] = y > x ? make_uint2(x, y) : make_uint2(0u, 0u);
    atomicAdd(&s_bk[bucket(y - x)], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (int b = 127; b >= 0; --b) {
      const uint32_t c = s_bk[b];
      s_bk[b] = run;
      run += c;
    }
  }
  __syncthreads();
  for (int t = tid; t < T; t += SCAN_THREADS) {
    const uint32_t x = __ldcg(starts + (size_t)t * S), y = __ldcg(starts + (size_t)(t + 1) * S);
    order[atomicAdd(&s_bk[bucket(y - x)], 1u)] = (uint32_t)t;
  }
}

// Exclusive scan of the n = T_total * S sub-bin counters into starts[0..n] (clamped to `capacity`),
// R = starts[n] unclamped into hdr[1].  Single pass over ceil(n / 4096) blocks, chained with the
// decoupled look-back; a block's place in the chain is its ticket (hdr[0]), not blockIdx, so the
// look-back never waits for a block that has not started.  state[b] = flag << 32 | value with
// flag 1 = block aggregate, 2 = inclusive prefix (one 64-bit word: relaxed accesses suffice).
// The block that FINISHES last (counter hdr[2]) turns the starts into `ranges` and the blend
// launch order — the work of a second kernel without its launch.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_starts_kernel(const uint32_t* __restrict__ counts, int n, uint32_t capacity,
                   uint32_t* starts, uint32_t* __restrict__ hdr,
                   unsigned long long* __restrict__ state, uint32_t* __restrict__ host_R,
                   int T, int S, uint2* __restrict__ ranges, uint32_t* __restrict__ order) {
  __shared__ uint32_t s_warp[SCAN_THREADS / 32];
  __shared__ uint32_t s_bid, s_prefix, s_last, s_bk[128];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_bid = atomicAdd(hdr, 1u);
  __syncthreads();
  const uint32_t bid = s_bid;
  const int i0 = (int)bid * SCAN_TILE + (int)tid * SCAN_ITEMS;
  uint32_t c[SCAN_ITEMS];
  if (i0 + SCAN_ITEMS <= n) {
    const uint4 q = *reinterpret_cast<const uint4*>(counts + i0);
    c[0] = q.x; c[1] = q.y; c[2] = q.z; c[3] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) c[k] = (i0 + k < n) ? counts[i0 + k] : 0u;
  }
  const uint32_t mine = c[0] + c[1] + c[2] + c[3];
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    const uint32_t t = s_warp[w];
    if ((uint32_t)w < warp) wbase += t;
    total += t;
  }
  if (warp == 0) {
    uint32_t excl = 0;
    if (bid == 0) {
      if (lane == 0) st_release_u64(state, (2ull << 32) | total);
    } else {
      if (lane == 0) st_release_u64(state + bid, (1ull << 32) | total);
      int base = (int)bid - 1;
      while (true) {
        const int j = base - (int)lane;
        unsigned long long sv = 0;
        if (j >= 0) {
          do { sv = ld_acquire_u64(state + j); } while ((sv >> 32) == 0ull);
        }
        const uint32_t flag = j >= 0 ? (uint32_t)(sv >> 32) : 2u;   // virtual prefix 0 before block 0
        const uint32_t val = j >= 0 ? (uint32_t)sv : 0u;
        const uint32_t pmask = __ballot_sync(0xffffffffu, flag == 2u);
        const int first = __ffs(pmask) - 1;       // nearest predecessor that already owns a prefix
        uint32_t v = (pmask == 0u || lane <= (uint32_t)first) ? val : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (pmask) break;
        base -= 32;
      }
      if (lane == 0) st_release_u64(state + bid, (2ull << 32) | (unsigned long long)(excl + total));
    }
    if (lane == 0) s_prefix = excl;
  }
  __syncthreads();
  uint32_t x = s_prefix + wbase + incl - mine;
  uint32_t o4[SCAN_ITEMS];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { o4[k] = min(x, capacity); x += c[k]; }
  if (i0 + SCAN_ITEMS <= n) {
    *reinterpret_cast<uint4*>(starts + i0) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) if (i0 + k < n) starts[i0 + k] = o4[k];
  }
  if (bid == gridDim.x - 1 && tid == 0) {
    const uint32_t R = s_prefix + total;
    starts[n] = min(R, capacity);
    hdr[1] = R;   // num_rendered
    // ... and straight into the caller's mapped pinned word: the host learns R from this store
    // plus an event, with no copy-engine transfer behind whatever bulk copies other streams queued
    if (host_R) { *reinterpret_cast<volatile uint32_t*>(host_R) = R; __threadfence_system(); }
  }
  __threadfence();                         // this block's starts are visible before it counts as done
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(hdr + 2, 1u) == gridDim.x - 1u ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  tile_ranges_and_order(starts, T, S, ranges, order, s_bk);
}

// grid: (ceil(P/256), nviews).  Same tile enumeration as the counting in the preprocess kernel.
__global__ void __launch_bounds__(GFT_BLOCK)
scatter_entries_kernel(const __grid_constant__ PreprocessParams p, const uint32_t* __restrict__ starts,
                       uint32_t* __restrict__ cursors, u64* __restrict__ entries, uint32_t capacity) {
  const int v = blockIdx.y;
  const ViewCam& vc = p.views[v];
  const int idx = (int)(blockIdx.x * GFT_BLOCK + threadIdx.x);
  const uint32_t lane = threadIdx.x & 31;
  const size_t P = (size_t)p.P;
  uint32_t rx0 = 0, ry0 = 0, rx1 = 0, tiles = 0;
  if (idx < p.P && vc.radii[idx] > 0) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p.g.rect) + v * P + idx);
    rx0 = r.x & 0xffffu; ry0 = r.x >> 16; rx1 = r.y & 0xffffu;
    const uint32_t ry1 = r.y >> 16;
    tiles = (rx1 - rx0) * (ry1 - ry0);
  }
  const uint32_t gx = (uint32_t)vc.grid_x, tb = (uint32_t)vc.tile_base;
  const uint32_t S = (uint32_t)p.sub_bins;
  const uint32_t zbits = tiles ? __float_as_uint(__ldg(p.g.depths + v * P + idx)) : 0u;
  // slot = start of the (tile, sub-bin) + the cursor's old value.  The starts are clamped to
  // `capacity` by the scan, so slot >= capacity can only happen when a caller's size hint was too
  // small (the call is then redone with the exact size); otherwise slot < starts[i + 1] holds by
  // construction (a sub-bin hands out exactly as many slots as the preprocess counted).
  for_each_tile_2phase(
      rx0, ry0, rx1, tiles, (uint32_t)idx, zbits, lane,
      [&](uint32_t tx, uint32_t ty, uint32_t g) {
        const uint32_t i = (tb + ty * gx + tx) * S + (g & (S - 1u));
        return make_uint2(__ldg(starts + i), atomicAdd(cursors + i, 1u));
      },
      [&](uint32_t g, uint32_t z, uint2 h) {
        const uint32_t slot = h.x + h.y;
        if (slot < capacity) entries[slot] = ((u64)z << 32) | (u64)g;
      });
}

// ---- per-segment sort ----------------------------------------------------------------------
// Bitonic sorting network in its "flip" form: level k first compares i with its mirror inside the
// k-block (i ^ (k-1)), then runs half-cleaners of distance k/4, k/8, ..., 1.  Every comparison
// puts the smaller key at the lower index, so a segment of arbitrary length n needs no padding:
// elements past n behave as +infinity and a comparison whose upper index is >= n is a no-op.
// Segments of up to `cap` entries (a power of two, the dynamic shared memory size / 8) are sorted
// entirely in shared memory; longer ones (initialisation-like clouds: thousands of large splats
// per tile) sort cap-sized chunks in shared memory and run the stages of distance >= cap on the
// segment in global memory (L2-resident), one block per segment either way.
constexpr int SORT_THREADS = 256;

__device__ __forceinline__ void ce_shared(u64* s, uint32_t a, uint32_t b) {
  const u64 x = s[a], y = s[b];
  if (x > y) { s[a] = y; s[b] = x; }
}

// all stages of distance < min(cap, k) of level k on the shared chunk s[0, len); k and every
// distance j are powers of two, so the index arithmetic is shifts and masks (a division by a
// run-time value costs ~25 instructions and made this network instruction-bound)
__device__ __forceinline__ void level_in_shared(u64* s, uint32_t len, uint32_t k, bool with_flip,
                                                uint32_t pairs) {
  if (with_flip) {
    const uint32_t half = k >> 1, lh = 31u - __clz(half);
    for (uint32_t t = threadIdx.x; t < pairs; t += SORT_THREADS) {
      const uint32_t lo = t & (half - 1), base = (t >> lh) << (lh + 1);
      const uint32_t a = base + lo, b = base + (k - 1u - lo);
      if (b < len) ce_shared(s, a, b);
    }
    __syncthreads();
  }
  for (uint32_t j = with_flip ? (k >> 2) : (k >> 1); j > 0; j >>= 1) {
    const uint32_t lj = 31u - __clz(j);
    for (uint32_t t = threadIdx.x; t < pairs; t += SORT_THREADS) {
      const uint32_t a = ((t >> lj) << (lj + 1)) + (t & (j - 1)), b = a + j;
      if (b < len) ce_shared(s, a, b);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t next_pow2(uint32_t n) {
  return n <= 1u ? 1u : (1u << (32 - __clz(n - 1u)));
}

// Stable LSD radix sort of s_a[0,n) on `passes` 8-bit digits of (depth_bits - kmin) >> sh0, ping-pong
// between s_a and s_b; returns the buffer that holds the result.  kmin is the segment's smallest
// depth word, so only the bits in which the segment's depths actually differ are sorted: at most
// the top 24 of them (3 passes; 2 or 1 when the depths of a tile span fewer bits) — whatever
// order the remaining low bits and the index ties need is restored by the caller's fix-up.
// 256 threads = 256 digit bins.  Warp w owns the contiguous span [w*span, (w+1)*span) and walks it
// 32 entries at a time, so ranks follow the current order (stability): per pass a per-warp digit
// histogram (shared-memory atomics), an exclusive scan over (digit, warp), then every 32-entry
// row is ranked by digit matching against the warp's running digit offsets and scattered into
// the other buffer.
template <bool HW_MATCH>
__device__ __forceinline__ u64* radix_depth_sort_shared(u64* s_a, u64* s_b, uint32_t* s_hist,
                                                        uint32_t* s_wtot, uint32_t n, uint32_t kmin,
                                                        uint32_t sh0, int passes) {
  constexpr int WARPS = SORT_THREADS / 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t items = (n + SORT_THREADS - 1) / SORT_THREADS;
  const uint32_t span = items * 32u;
  const uint32_t w0 = warp * span;
  u64* src = s_a;
  u64* dst = s_b;
  for (int pass = 0; pass < passes; ++pass) {
    const uint32_t shift = sh0 + 8u * (uint32_t)pass;
    for (uint32_t i = tid; i < WARPS * 256u; i += SORT_THREADS) s_hist[i] = 0u;
    __syncthreads();
    for (uint32_t i = 0; i < items; ++i) {
      const uint32_t e = w0 + i * 32u + lane;
      if (e < n) atomicAdd(&s_hist[warp * 256u + ((((uint32_t)(src[e] >> 32) - kmin) >> shift) & 0xffu)], 1u);
    }
    __syncthreads();
    {
      // digit d = tid: exclusive prefix over warps, then over digits
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const uint32_t c = s_hist[w * 256 + tid];
        s_hist[w * 256 + tid] = run;
        run += c;
      }
      uint32_t incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
      }
      if (lane == 31) s_wtot[warp] = incl;
      __syncthreads();
      uint32_t wb = 0;
      for (uint32_t w = 0; w < warp; ++w) wb += s_wtot[w];
      const uint32_t excl = wb + incl - run;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) s_hist[w * 256 + tid] += excl;
    }
    __syncthreads();
    for (uint32_t i = 0; i < items; ++i) {
      const uint32_t e = w0 + i * 32u + lane;
      const bool ok = e < n;
      const u64 key = ok ? src[e] : 0ull;
      const uint32_t d = ok ? ((((uint32_t)(key >> 32) - kmin) >> shift) & 0xffu) : 0x100u + lane;   // padding lanes match nobody
      const uint32_t peers = HW_MATCH ? __match_any_sync(0xffffffffu, d) : match_digit8(d, ok);
      const uint32_t leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader && ok) {
        old = s_hist[warp * 256u + d];
        s_hist[warp * 256u + d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, ok ? leader : lane);
      if (ok) dst[old + __popc(peers & ((1u << lane) - 1u))] = key;
      __syncwarp();
    }
    __syncthreads();
    u64* t = src; src = dst; dst = t;
  }
  return src;
}

__global__ void __launch_bounds__(SORT_THREADS)
tile_sort_kernel(const uint2* __restrict__ ranges, u64* __restrict__ entries,
                 uint32_t* __restrict__ point_list, uint32_t cap, int use_radix, int adapt) {
  extern __shared__ __align__(16) unsigned char sort_smem_raw[];
  u64* s = reinterpret_cast<u64*>(sort_smem_raw);
  __shared__ uint32_t s_wtot[SORT_THREADS / 32], s_wmax[SORT_THREADS / 32];
  const uint2 rg = ranges[blockIdx.x];
  const uint32_t n = rg.y - rg.x;
  if (n == 0u) return;
  u64* g = entries + rg.x;
  uint32_t* pl = point_list + rg.x;
  const uint32_t tid = threadIdx.x;

  if (n <= cap) {
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    uint32_t zlo = 0xffffffffu, zhi = 0u;
    for (uint32_t i = tid; i < n; i += SORT_THREADS) {
      const u64 e = g[i];
      s[i] = e;
      zlo = min(zlo, (uint32_t)(e >> 32));
      zhi = max(zhi, (uint32_t)(e >> 32));
    }
    u64* cur = s;
    bool unsorted = true, try_fixup = false;
    const bool radix = use_radix && n > 32u;
    if (radix) {
      zlo = __reduce_min_sync(0xffffffffu, zlo);
      zhi = __reduce_max_sync(0xffffffffu, zhi);
      if (lane == 0) { s_wtot[warp] = zlo; s_wmax[warp] = zhi; }
    }
    __syncthreads();                                       // the segment (and the warps' min / max) are in shared memory
    if (radix) {
#pragma unroll
      for (int w = 0; w < SORT_THREADS / 32; ++w) { zlo = min(zlo, s_wtot[w]); zhi = max(zhi, s_wmax[w]); }
      __syncthreads();                                     // s_wtot is reused by the radix passes
      // the depth words of this tile differ in their low `bits` bits only (after subtracting the
      // smallest): sort on the top min(bits, 24) of those
      if (!adapt) { zlo = 0u; zhi = 0xffffffffu; }         // A/B: all 32 bits in four passes
      const uint32_t bits = 32u - (uint32_t)__clz(zhi - zlo);
      const int passes = adapt ? (int)min(3u, (bits + 7u) >> 3) : 4;
      const uint32_t sh0 = (adapt && bits > 24u) ? bits - 24u : 0u;
      u64* s_b = s + cap;
      uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_b + cap);
      cur = use_radix == 2 ? radix_depth_sort_shared<true>(s, s_b, s_hist, s_wtot, n, zlo, sh0, passes)
                           : radix_depth_sort_shared<false>(s, s_b, s_hist, s_wtot, n, zlo, sh0, passes);
      // What is left: entries that agree in the sorted bits (a few low depth bits when the tile's
      // depths span more than 24 bits; equal depths, whose slots came from atomics in any order)
      // may be the wrong way round.  Such runs are short, so an odd-even transposition on the
      // whole 64-bit entries restores (depth bits, index) order in one or two rounds; only a tile
      // that still moves after 8 rounds takes the bitonic network.
      int bad = 0;
      for (uint32_t i = tid; i + 1 < n; i += SORT_THREADS) bad |= (cur[i] > cur[i + 1]) ? 1 : 0;
      unsorted = __syncthreads_or(bad) != 0;
      try_fixup = unsorted;
    }
    for (int round = 0; try_fixup && unsorted && round < 8; ++round) {
      int moved = 0;
      for (uint32_t a = 2u * tid; a + 1u < n; a += 2u * SORT_THREADS) {
        const u64 x = cur[a], y = cur[a + 1];
        if (x > y) { cur[a] = y; cur[a + 1] = x; moved = 1; }
      }
      __syncthreads();
      for (uint32_t a = 2u * tid + 1u; a + 1u < n; a += 2u * SORT_THREADS) {
        const u64 x = cur[a], y = cur[a + 1];
        if (x > y) { cur[a] = y; cur[a + 1] = x; moved = 1; }
      }
      unsorted = __syncthreads_or(moved) != 0;
    }
    if (unsorted) {
      const uint32_t m = next_pow2(n);
      for (uint32_t k = 2; k <= m; k <<= 1) level_in_shared(cur, n, k, true, m >> 1);
    }
    for (uint32_t i = tid; i < n; i += SORT_THREADS) {
      const u64 e = cur[i];
      g[i] = e;
      pl[i] = (uint32_t)e;
    }
    return;
  }

  // ---- long segment: chunks in shared memory, wide stages in global memory --------------------
  const uint32_t m = next_pow2(n);
  for (uint32_t c0 = 0; c0 < n; c0 += cap) {
    const uint32_t len = min(cap, n - c0);
    for (uint32_t i = tid; i < len; i += SORT_THREADS) s[i] = __ldcg(g + c0 + i);
    __syncthreads();
    for (uint32_t k = 2; k <= cap; k <<= 1) level_in_shared(s, len, k, true, cap >> 1);
    for (uint32_t i = tid; i < len; i += SORT_THREADS) g[c0 + i] = s[i];
    __syncthreads();
  }
  for (uint32_t k = cap << 1; k <= m; k <<= 1) {
    const uint32_t half = k >> 1, lh = 31u - __clz(half);
    for (uint32_t t = tid; t < (m >> 1); t += SORT_THREADS) {       // flip, distance up to k-1
      const uint32_t lo = t & (half - 1), base = (t >> lh) << (lh + 1);
      const uint32_t a = base + lo, b = base + (k - 1u - lo);
      if (b < n) {
        const u64 x = __ldcg(g + a), y = __ldcg(g + b);
        if (x > y) { g[a] = y; g[b] = x; }
      }
    }
    __syncthreads();
    for (uint32_t j = k >> 2; j >= cap; j >>= 1) {                    // half-cleaners across chunks
      const uint32_t lj = 31u - __clz(j);
      for (uint32_t t = tid; t < (m >> 1); t += SORT_THREADS) {
        const uint32_t a = ((t >> lj) << (lj + 1)) + (t & (j - 1)), b = a + j;
        if (b < n) {
          const u64 x = __ldcg(g + a), y = __ldcg(g + b);
          if (x > y) { g[a] = y; g[b] = x; }
        }
      }
      __syncthreads();
    }
    for (uint32_t c0 = 0; c0 < n; c0 += cap) {                        // distances < cap: per chunk
      const uint32_t len = min(cap, n - c0);
      for (uint32_t i = tid; i < len; i += SORT_THREADS) s[i] = __ldcg(g + c0 + i);
      __syncthreads();
      level_in_shared(s, len, cap, false, cap >> 1);
      for (uint32_t i = tid; i < len; i += SORT_THREADS) g[c0 + i] = s[i];
      __syncthreads();
    }
  }
  for (uint32_t i = tid; i < n; i += SORT_THREADS) pl[i] = (uint32_t)__ldcg(g + i);
}

}  // namespace

void launch_tile_scan(const uint32_t* tile_counts, int T_total, int sub_bins, uint32_t capacity,
                      uint32_t* starts, uint2* ranges, uint32_t* order, uint32_t* hdr,
                      unsigned long long* scan_state, uint32_t* host_R, cudaStream_t stream) {
  const int n = T_total * sub_bins;
  scan_starts_kernel<<<(n + SCAN_TILE - 1) / SCAN_TILE, SCAN_THREADS, 0, stream>>>(tile_counts, n, capacity, starts,
                                                                                  hdr, scan_state, host_R, T_total, sub_bins, ranges,
                                                                                  order);
  note_launches(1);
}

void launch_scatter_entries(const PreprocessParams& pp, const uint32_t* starts, uint32_t* cursors,
                            unsigned long long* entries, uint32_t capacity, cudaStream_t stream) {
  if (pp.P <= 0) return;
  const dim3 grid((pp.P + GFT_BLOCK - 1) / GFT_BLOCK, pp.nviews);
  scatter_entries_kernel<<<grid, GFT_BLOCK, 0, stream>>>(pp, starts, cursors, entries, capacity);
  note_launches(1);
}

void launch_tile_sort(const uint2* ranges, int T_total, unsigned long long* entries,
                      uint32_t* point_list, int mean_len_hint, cudaStream_t stream) {
  if (T_total <= 0) return;
  // Entries per block held in shared memory.  Radix path: two buffers + 8 KB of histograms, i.e.
  // 40 KB at 2048 (5 blocks/SM), 72 KB at 4096 (3 blocks/SM); longer segments take the bitonic
  // path on 4096-entry chunks.  Option sort_cap forces a capacity, sort_radix = 0 the bitonic
  // network everywhere (A/B runs, tests of the long-segment path).
  const int forced = option(OPT_SORT_CAP);
  const int use_radix = option(OPT_SORT_RADIX) != 0 ? (option(OPT_SORT_MATCH) != 0 ? 2 : 1) : 0;
  uint32_t cap = mean_len_hint > 1400 ? 4096u : 2048u;
  if (forced == 256 || forced == 1024 || forced == 2048 || forced == 4096 || forced == 8192) cap = (uint32_t)forced;
  const int smem = use_radix ? (int)cap * 16 + 8 * 256 * 4 : (int)cap * 8;
  static unsigned long long smem_ok = 0;
  ensure_dynamic_smem(tile_sort_kernel, 8192 * 16 + 8 * 256 * 4, &smem_ok);
  tile_sort_kernel<<<T_total, SORT_THREADS, smem, stream>>>(ranges, entries, point_list, cap, use_radix,
                                                            option(OPT_SORT_ADAPT) != 0 ? 1 : 0);
  note_launches(1);
}

// ---- radix sort pass-throughs (knn, A/B hooks) -------------------------------------------------
}  // namespace gft

#include "radix_sort.cuh"

namespace gft {

bool sort_result_in_out(int end_bit) { return sort_lands_in_out(end_bit); }

size_t sort_pairs_temp_bytes(int R) { return radix_sort_temp_bytes(R); }

int sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
               const uint32_t* vals_in, uint32_t* vals_out, int R, int end_bit,
               cudaStream_t stream, const uint32_t* d_R) {
  return radix_sort_pairs(d_temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, R, end_bit,
                          stream, d_R);
}

}  // namespace gft
