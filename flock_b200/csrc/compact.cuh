// compact.cuh -- the single-pass stable compaction skeleton shared by FilterExec (filter_project.cu),
// the hash-table emitter (hash_agg.cu) and the partition selector (partition.cu).
//
// A persistent CTA takes tiles of CP_THREADS x I items from an atomic ticket.  For each tile it gets one
// bit per item ("survives"), ranks the survivors in item order with warp ballots, obtains the tile's
// global output offset by decoupled look-back (device_utils.cuh) and hands every survivor its output
// position.  The last CTA to finish resets the scratch, so back-to-back launches need no memset.
//
// I = items per thread per tile (16 / 32 / 64): measured on B200 (profiles/r1_microbench.txt), a 40 MB
// column is read fastest with >= 4 independent 16-byte loads in flight per thread and as FEW
// inter-tile synchronisation points as possible -- with I = 64 a 10 M-row relation is 611 tiles, fewer
// than the resident CTAs, so every CTA does exactly one ticket, one look-back and three barriers.
#pragma once

#include "device_utils.cuh"

namespace fg {

constexpr int CP_THREADS = 256;
constexpr int CP_WARPS = CP_THREADS / 32;

struct CompactScratch {
  unsigned long long* tile_state;  // look-back words, zero between launches
  unsigned int* counters;          // [0] ticket, [1] done
  unsigned long long* out_count;   // receives the total number of survivors
  long long num_tiles;
  int stride;         // u64 words between the look-back words of consecutive tiles (1 = packed; 32 = one 256-byte L2 chunk each)
  int poll_sleep_ns;  // back-off of a polling thread that found a predecessor not yet published (0 = spin)
};

// Item k of thread `tid` is item index  ((k / E) * CP_THREADS + tid) * E + k % E  of the tile: E = items
// a thread loads contiguously (vector width).  Survivors are ordered (group = k / E, thread, k % E),
// which is ascending item index.
template <int E>
__device__ __forceinline__ long long cp_item_index(int k, int tid) {
  return ((long long)(k / E) * CP_THREADS + tid) * E + (k % E);
}

template <int E, int I>
struct CompactSmem {
  static constexpr int G = I / E;
  unsigned group_warp[G][CP_WARPS];
  unsigned long long lb_sum[CP_WARPS];
  int lb_has[CP_WARPS];
  int lb_done;
  long long tile;
  unsigned long long excl;
  unsigned tile_total;
  int last;
};

// Fetches the next tile index for the CTA (or -1 when the work is exhausted).
template <int E, int I>
__device__ __forceinline__ long long cp_next_tile(CompactSmem<E, I>& s, const CompactScratch& sc) {
  if (threadIdx.x == 0) s.tile = (long long)atomicAdd(sc.counters, 1u);
  __syncthreads();
  long long t = s.tile;
  return t < sc.num_tiles ? t : -1;
}

// Decoupled look-back with the WHOLE CTA: thread i inspects tile (base - i), so one step covers CP_THREADS
// predecessors.  When a launch is a single wave (all tiles finish counting at about the same time) nearly every
// predecessor still shows PARTIAL, and a 32-wide window would walk back serially, one L2 round trip per 32 tiles
// (measured: 611 tiles -> ~19 dependent round trips, the dominant cost of the first version, see
// profiles/r1_filter_ncu.md); 256-wide windows need at most ceil(tiles / 256) steps.  Leaves the exclusive
// prefix of `tile` in s.excl (valid for every thread after the trailing barrier).
template <int E, int I>
__device__ __forceinline__ void cp_block_lookback(CompactSmem<E, I>& s, const CompactScratch& sc, long long tile, unsigned total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long excl = 0;  // meaningful in thread 0 only
  if (tile == 0) {
    if (tid == 0) st_relaxed_u64(sc.tile_state, LB_PREFIX | (unsigned long long)total);
  } else {
    if (tid == 0) st_relaxed_u64(sc.tile_state + tile * sc.stride, LB_PARTIAL | (unsigned long long)total);
    long long base = tile - 1;
    while (true) {
      const long long t = base - tid;
      unsigned long long st = LB_PREFIX;  // virtual tiles before the first one: inclusive prefix 0
      if (t >= 0) {
        while (true) {
          st = ld_relaxed_u64(sc.tile_state + t * sc.stride);
          if ((st & LB_FLAG_MASK) != LB_INVALID) break;
          if (sc.poll_sleep_ns) __nanosleep(sc.poll_sleep_ns);
        }
      }
      const unsigned pm = __ballot_sync(FULL_MASK, (st & LB_FLAG_MASK) == LB_PREFIX);
      const int first = pm ? __ffs(pm) - 1 : 32;
      const unsigned long long wsum = warp_sum(lane <= first ? (st & LB_VALUE_MASK) : 0ull);
      if (lane == 0) {
        s.lb_sum[warp] = wsum;
        s.lb_has[warp] = pm != 0;
      }
      __syncthreads();
      if (tid == 0) {
        int done = 0;
        for (int w = 0; w < CP_WARPS && !done; ++w) {  // warp 0 holds the nearest predecessors
          excl += s.lb_sum[w];
          done = s.lb_has[w];
        }
        s.lb_done = done;
      }
      __syncthreads();
      if (s.lb_done) break;
      base -= CP_THREADS;
    }
    if (tid == 0) st_relaxed_u64(sc.tile_state + tile * sc.stride, LB_PREFIX | (excl + total));
  }
  if (tid == 0) {
    s.excl = excl;
    if (tile == sc.num_tiles - 1) *sc.out_count = excl + total;
  }
  __syncthreads();
}

// Ranks the survivors of one tile (bit k of `bits` = item k survives).  On return (after the internal
// barriers) cp_position() gives every survivor its output slot and s.tile_total the tile's survivor count.
template <int E, int I>
__device__ __forceinline__ void cp_rank_tile(CompactSmem<E, I>& s, const CompactScratch& sc, long long tile, unsigned long long bits,
                                             unsigned (&lane_prefix)[I / E]) {
  constexpr int G = I / E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int g = 0; g < G; ++g) {
    unsigned pre = 0, tot = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      unsigned b = __ballot_sync(FULL_MASK, (bits >> (g * E + e)) & 1ull);
      pre += __popc(b & lt);
      tot += __popc(b);
    }
    lane_prefix[g] = pre;
    if (lane == 0) s.group_warp[g][warp] = tot;
  }
  __syncthreads();
  if (warp == 0) {
    constexpr int N = G * CP_WARPS;
    constexpr int PER = (N + 31) / 32;
    unsigned* flat = &s.group_warp[0][0];
    unsigned v[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int idx = lane * PER + i;
      v[i] = idx < N ? flat[idx] : 0u;
      sum += v[i];
    }
    unsigned incl = warp_inclusive_sum(sum);
    unsigned run = incl - sum;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int idx = lane * PER + i;
      if (idx < N) flat[idx] = run;
      run += v[i];
    }
    unsigned total = __shfl_sync(FULL_MASK, incl, 31);
    if (lane == 0) s.tile_total = total;
  }
  __syncthreads();
  cp_block_lookback(s, sc, tile, s.tile_total);
}

template <int E, int I>
__device__ __forceinline__ long long cp_position(const CompactSmem<E, I>& s, unsigned long long bits, int k, const unsigned (&lane_prefix)[I / E]) {
  const int g = k / E, e = k % E;
  const unsigned within = __popcll(bits & (((1ull << e) - 1ull) << (g * E)));
  return (long long)s.excl + s.group_warp[g][threadIdx.x >> 5] + lane_prefix[g] + within;
}

// Call once per CTA after its tile loop: the last CTA to arrive clears the scratch.
template <int E, int I>
__device__ __forceinline__ void cp_finish(CompactSmem<E, I>& s, const CompactScratch& sc) {
  if (threadIdx.x == 0) {
    __threadfence();  // one thread: a CTA-wide fence costs ~1.5 us of L1 invalidation on every SM (profiles/r1_filter_ncu.md)
    s.last = (atomicAdd(sc.counters + 1, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s.last) {
    for (long long i = threadIdx.x; i < sc.num_tiles; i += CP_THREADS) sc.tile_state[i * sc.stride] = LB_INVALID;
    if (threadIdx.x == 0) {
      sc.counters[0] = 0;
      sc.counters[1] = 0;
    }
  }
}

}  // namespace fg
