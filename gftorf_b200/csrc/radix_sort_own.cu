// radix_sort_own.cu — hand-written stable LSD radix sort for (u64 key, u32 value) pairs, sm_100a.
//
// Shape: 8-bit digits; ONE histogram kernel reads the keys once and builds the digit histograms
// of every pass; each pass is then ONE "onesweep" kernel: a tile of TILE_ITEMS consecutive pairs
// is ranked inside the block (warp match_any ranking, stable), the tile's digit counts are
// chained to the preceding tiles with a decoupled look-back (so the global digit offsets never
// need a separate scan kernel or a second read of the keys), pairs are reordered by digit in
// shared memory and written out in runs so that stores coalesce.
//
// HBM traffic: 8 B/pair for the histogram + (12 B read + 12 B write)/pair/pass — the same
// algorithmic traffic as cub's onesweep; what it saves against the reference's call is the
// temp-buffer copy of the non-double-buffered CUB API and the separate upsweep launches.
#include "common.cuh"
#include "radix_sort.cuh"
#include "kernels.h"

#include <cstdlib>

namespace gft {

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
// pairs per thread: 16 (4096-pair tiles).  An 8-pair variant (2048-pair tiles, 64 registers, four
// resident blocks per SM) is kept behind GFT_SORT_ITEMS=8: measured on B200 it is 5 % SLOWER at
// R ~ 1e6 and equal at R ~ 1e7 — a pass there is bound by its fixed costs (launch, ramp, the
// load -> rank -> prefix -> reorder -> store dependency chain), not by resident warps.
constexpr int RS_ITEMS_MAX = 16;
constexpr int RS_ITEMS_MIN = 8;
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 8;

constexpr int RS_LOOKBACK = 8;
constexpr uint32_t RS_GROUP = 16;   // tiles per look-back group
constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PRE = 2u << 30;
constexpr uint32_t VAL_MASK = (1u << 30) - 1;

// The look-back status word carries its payload (flag | count) in the same 32 bits, so relaxed
// gpu-scope accesses are enough: nothing else has to be ordered after the flag.  (An acquire load
// costs an L1 invalidate, CCTL.IVALL, on every poll.)
__device__ __forceinline__ uint32_t ld_status_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// hist[pass][bin] over all keys; one read of the keys for all passes.
__global__ void __launch_bounds__(RS_THREADS)
rs_histogram_kernel(const uint64_t* __restrict__ keys, int R_cap, const uint32_t* __restrict__ d_R,
                    int npass, int end_bit, uint32_t* __restrict__ hist) {
  // the pair count lives on the device when the caller sized the buffers from a hint
  const int R = d_R ? (int)min(__ldg(d_R), (uint32_t)R_cap) : R_cap;
  __shared__ uint32_t s_hist[RS_MAX_PASSES * RS_BINS];
  for (int i = threadIdx.x; i < npass * RS_BINS; i += RS_THREADS) s_hist[i] = 0;
  __syncthreads();
  const int stride = gridDim.x * RS_THREADS;
  for (int i = blockIdx.x * RS_THREADS + threadIdx.x; i < R; i += stride) {
    const uint64_t k = __ldg(keys + i);
    for (int p = 0; p < npass; ++p) {
      const int shift = 8 * p;
      const int nb = min(8, end_bit - shift);
      const uint32_t d = (uint32_t)(k >> shift) & ((1u << nb) - 1u);
      atomicAdd(&s_hist[p * RS_BINS + d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * RS_BINS; i += RS_THREADS) {
    const uint32_t v = s_hist[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// In-place exclusive scan of each pass's 256 bins: hist -> global digit base.
__global__ void __launch_bounds__(RS_BINS) rs_scan_bins_kernel(uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_w[RS_BINS / 32];
  uint32_t* h = hist + blockIdx.x * RS_BINS;
  const uint32_t v = h[threadIdx.x];
  uint32_t incl = v;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += n;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; ++w) base += s_w[w];
  h[threadIdx.x] = base + incl - v;
}

template <int RS_TILE>
struct RsSmemT {
  uint64_t keys[RS_TILE];
  uint32_t vals[RS_TILE];
  uint32_t warp_hist[RS_WARPS][RS_BINS];  // per-warp digit counts -> exclusive warp prefixes
  uint32_t tile_excl[RS_BINS];            // exclusive prefix of digit counts inside the tile
  uint32_t gbase[RS_BINS];                // global output position of the tile's first digit-d pair
  uint32_t tile_id;
};

template <int RS_ITEMS, int MINB>
__global__ void __launch_bounds__(RS_THREADS, MINB)
rs_onesweep_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                   const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, int R_cap,
                   const uint32_t* __restrict__ d_R, int shift, int nbits,
                   const uint32_t* __restrict__ digit_base, uint32_t* __restrict__ ticket,
                   uint32_t* __restrict__ state, uint32_t* __restrict__ gstate) {
  constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
  using RsSmem = RsSmemT<RS_TILE>;
  extern __shared__ __align__(16) unsigned char rs_smem_raw[];
  RsSmem& s = *reinterpret_cast<RsSmem*>(rs_smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = d_R ? (int)min(__ldg(d_R), (uint32_t)R_cap) : R_cap;

  if (tid == 0) s.tile_id = atomicAdd(ticket, 1u);
  for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&s.warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s.tile_id;
  const uint32_t tile_base = tile * RS_TILE;
  // the grid is sized for R_cap; tiles past the actual count have nothing to do and nobody waits
  // on them (tickets are handed out in order, so every real tile is claimed by some block)
  if (tile_base >= (uint32_t)R) return;
  const uint32_t dmask = (1u << nbits) - 1u;

  // ---- load (warp-striped: warp w owns RS_ITEMS*32 consecutive pairs) and rank -------------
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint32_t rnk[RS_ITEMS];  // rank among the warp's pairs with the same digit
  const uint32_t wbase = tile_base + warp * (RS_ITEMS * 32);
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const uint32_t g = wbase + i * 32 + lane;
    const bool ok = g < (uint32_t)R;
    key[i] = ok ? __ldg(keys_in + g) : ~0ull;
    val[i] = ok ? __ldg(vals_in + g) : 0u;
  }
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const uint32_t g = wbase + i * 32 + lane;
    const bool ok = g < (uint32_t)R;
    const uint32_t d = ok ? ((uint32_t)(key[i] >> shift) & dmask) : 256u;  // 256 = padding
    // eight ballots instead of MATCH.ANY (common.cuh); padding lanes form their own group either
    // way and never touch the histogram
    const uint32_t peers = match_digit8(d, ok) | (ok ? 0u : (1u << lane));
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && ok) {
      old = s.warp_hist[warp][d];
      s.warp_hist[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rnk[i] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();

  // ---- per-digit: prefix over warps, tile count, look-back, in-tile exclusive scan ---------
  {
    const uint32_t d = tid;  // RS_THREADS == RS_BINS
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = s.warp_hist[w][d];
      s.warp_hist[w][d] = run;
      run += c;
    }
    const uint32_t count = run;
    // ---- exclusive prefix of this tile's digit count over all earlier tiles -----------------
    // Two levels.  A plain decoupled look-back over tiles is a serial chain here: at R ~ 1e6 every
    // tile of the pass is resident at once, nobody owns an inclusive prefix yet, and prefixes
    // propagate at ~16 tiles per L2 round trip (measured: 13 of the 18 us of a pass).  So tiles are
    // grouped by RS_GROUP: inside a group a tile sums its predecessors' aggregates directly (one
    // round trip, <= RS_GROUP-1 independent loads); the LAST tile of a group publishes the group
    // total and runs the decoupled look-back over GROUPS (a 16x shorter chain); the other tiles
    // of a group wait for the previous group's inclusive prefix, a single word.
    const uint32_t tiles_real = ((uint32_t)R + RS_TILE - 1) / RS_TILE;
    const uint32_t grp = tile / RS_GROUP, first = grp * RS_GROUP;
    const uint32_t last = min(first + RS_GROUP - 1, tiles_real - 1);
    st_status_u32(state + (size_t)tile * RS_BINS + d, FLAG_AGG | count);
    uint32_t in_grp = 0;
    {
      uint32_t v[RS_GROUP - 1];
#pragma unroll
      for (int i = 0; i < RS_GROUP - 1; ++i)
        v[i] = (first + i < tile) ? ld_status_u32(state + (size_t)(first + i) * RS_BINS + d) : FLAG_AGG;
#pragma unroll
      for (int i = 0; i < RS_GROUP - 1; ++i) {
        while ((v[i] & ~VAL_MASK) == 0u) v[i] = ld_status_u32(state + (size_t)(first + i) * RS_BINS + d);
        in_grp += v[i] & VAL_MASK;
      }
    }
    uint32_t gexcl = 0;
    if (tile == last) {
      const uint32_t gtotal = in_grp + count;
      uint32_t* gs = gstate + (size_t)grp * RS_BINS + d;
      if (grp == 0) {
        st_status_u32(gs, FLAG_PRE | gtotal);
      } else {
        st_status_u32(gs, FLAG_AGG | gtotal);
        int t = (int)grp - 1;
        bool found = false;
        while (!found) {
          uint32_t v[RS_LOOKBACK];
#pragma unroll
          for (int i = 0; i < RS_LOOKBACK; ++i)
            v[i] = (t - i >= 0) ? ld_status_u32(gstate + (size_t)(t - i) * RS_BINS + d) : FLAG_PRE;
#pragma unroll
          for (int i = 0; i < RS_LOOKBACK; ++i) {
            if (!found) {
              while ((v[i] & ~VAL_MASK) == 0u) v[i] = ld_status_u32(gstate + (size_t)(t - i) * RS_BINS + d);
              gexcl += v[i] & VAL_MASK;
              if (v[i] & FLAG_PRE) found = true;
            }
          }
          t -= RS_LOOKBACK;
        }
        st_status_u32(gs, FLAG_PRE | (gexcl + gtotal));
      }
    } else if (grp != 0) {
      const uint32_t* gp = gstate + (size_t)(grp - 1) * RS_BINS + d;
      uint32_t v = ld_status_u32(gp);
      while ((v & FLAG_PRE) == 0u) v = ld_status_u32(gp);
      gexcl = v & VAL_MASK;
    }
    const uint32_t excl = gexcl + in_grp;
    // exclusive scan of `count` over the 256 digits
    uint32_t incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += n;
    }
    __shared__ uint32_t s_wtot[RS_WARPS];
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (uint32_t w = 0; w < warp; ++w) wb += s_wtot[w];
    const uint32_t texcl = wb + incl - count;
    s.tile_excl[d] = texcl;
    // global position of local slot i (digit d): gbase[d] + i, with gbase = base + excl - texcl
    s.gbase[d] = __ldg(digit_base + d) + excl - texcl;
  }
  __syncthreads();

  // ---- reorder by digit in shared memory ---------------------------------------------------
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const uint32_t g = wbase + i * 32 + lane;
    if (g < (uint32_t)R) {
      const uint32_t d = (uint32_t)(key[i] >> shift) & dmask;
      const uint32_t slot = s.tile_excl[d] + s.warp_hist[warp][d] + rnk[i];
      s.keys[slot] = key[i];
      s.vals[slot] = val[i];
    }
  }
  __syncthreads();

  // ---- coalesced write-out -----------------------------------------------------------------
  const uint32_t n_valid = min((uint32_t)RS_TILE, (uint32_t)R - tile_base);
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const uint32_t slot = i * RS_THREADS + tid;
    if (slot < n_valid) {
      const uint64_t k = s.keys[slot];
      const uint32_t d = (uint32_t)(k >> shift) & dmask;
      const uint32_t pos = s.gbase[d] + slot;
      keys_out[pos] = k;
      vals_out[pos] = s.vals[slot];
    }
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

int radix_sort_passes(int end_bit) { return (end_bit + 7) / 8; }

namespace {
constexpr int RS_TILE_MIN = RS_THREADS * RS_ITEMS_MIN;
constexpr int RS_TILE_MAX = RS_THREADS * RS_ITEMS_MAX;
inline int pick_items(int /*R*/) {
  static const int forced = [] {
    const char* e = std::getenv("GFT_SORT_ITEMS");
    return (e && e[0] == '8') ? RS_ITEMS_MIN : RS_ITEMS_MAX;
  }();
  return forced;
}
}  // namespace

size_t own_sort_temp_bytes(int R) {
  const size_t tiles = ((size_t)max(R, 1) + RS_TILE_MIN - 1) / RS_TILE_MIN;
  // hist[8][256] + ticket[8] + per pass: tile state [tiles][256] + group state [groups][256]
  const size_t groups = (tiles + RS_GROUP - 1) / RS_GROUP;
  return align_up(RS_MAX_PASSES * RS_BINS * 4 + RS_MAX_PASSES * 4, 256) +
         RS_MAX_PASSES * (tiles + groups) * RS_BINS * 4;
}

// Sorts on bits [0,end_bit).  Ping-pongs between (a) and (b) starting from a; the result lands in
// b when the pass count is odd and in a when it is even.  Returns the pass count, <0 on error.
int own_sort_pairs(void* d_temp, size_t temp_bytes, const uint64_t* keys_a_c, uint64_t* keys_b,
                   const uint32_t* vals_a_c, uint32_t* vals_b, int R, int end_bit,
                   cudaStream_t stream, const uint32_t* d_R) {
  if (R <= 0) return 0;
  if (end_bit > 64 || end_bit <= 0) return -1;
  // the look-back status words carry a 30-bit count next to the two flag bits
  if ((unsigned)R > VAL_MASK) return -4;
  if (temp_bytes < own_sort_temp_bytes(R)) return -2;
  uint64_t* keys_a = const_cast<uint64_t*>(keys_a_c);
  uint32_t* vals_a = const_cast<uint32_t*>(vals_a_c);
  const int npass = radix_sort_passes(end_bit);
  const int items = pick_items(R);
  const int tiles = (R + RS_THREADS * items - 1) / (RS_THREADS * items);
  const size_t head = align_up(RS_MAX_PASSES * RS_BINS * 4 + RS_MAX_PASSES * 4, 256);
  uint32_t* hist = reinterpret_cast<uint32_t*>(d_temp);
  uint32_t* ticket = hist + RS_MAX_PASSES * RS_BINS;
  uint32_t* state = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(d_temp) + head);
  const int groups = (tiles + (int)RS_GROUP - 1) / (int)RS_GROUP;
  const size_t per_pass = (size_t)(tiles + groups) * RS_BINS;
  cudaMemsetAsync(d_temp, 0, head + (size_t)npass * per_pass * 4, stream);

  int hblocks = (R + RS_THREADS * 8 - 1) / (RS_THREADS * 8);
  hblocks = max(1, min(hblocks, sm_count() * 8));
  rs_histogram_kernel<<<hblocks, RS_THREADS, 0, stream>>>(keys_a, R, d_R, npass, end_bit, hist);
  rs_scan_bins_kernel<<<npass, RS_BINS, 0, stream>>>(hist);
  note_launches(2 + npass);

  static unsigned long long smem_ok_min = 0, smem_ok_max = 0;
  if (items == RS_ITEMS_MIN)
    ensure_dynamic_smem(rs_onesweep_kernel<RS_ITEMS_MIN, 4>, (int)sizeof(RsSmemT<RS_TILE_MIN>), &smem_ok_min);
  else
    ensure_dynamic_smem(rs_onesweep_kernel<RS_ITEMS_MAX, 2>, (int)sizeof(RsSmemT<RS_TILE_MAX>), &smem_ok_max);
  uint64_t* kin = keys_a; uint64_t* kout = keys_b;
  uint32_t* vin = vals_a; uint32_t* vout = vals_b;
  for (int p = 0; p < npass; ++p) {
    const int shift = 8 * p;
    const int nb = min(8, end_bit - shift);
    uint32_t* st = state + (size_t)p * per_pass;
    if (items == RS_ITEMS_MIN)
      rs_onesweep_kernel<RS_ITEMS_MIN, 4><<<tiles, RS_THREADS, sizeof(RsSmemT<RS_TILE_MIN>), stream>>>(
          kin, kout, vin, vout, R, d_R, shift, nb, hist + p * RS_BINS, ticket + p, st, st + (size_t)tiles * RS_BINS);
    else
      rs_onesweep_kernel<RS_ITEMS_MAX, 2><<<tiles, RS_THREADS, sizeof(RsSmemT<RS_TILE_MAX>), stream>>>(
          kin, kout, vin, vout, R, d_R, shift, nb, hist + p * RS_BINS, ticket + p, st, st + (size_t)tiles * RS_BINS);
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  return npass;
}

}  // namespace gft
