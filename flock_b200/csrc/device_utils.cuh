// device_utils.cuh -- device-side helpers shared by the kernels: cache-hinted vector loads/stores,
// warp scans, the decoupled look-back tile-prefix protocol, hashing, TMA-bulk (cp.async.bulk) and
// mbarrier wrappers for sm_100a.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define FG_STR2(x) #x
#define FG_STR(x) FG_STR2(x)

namespace fg {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ---- streaming loads / stores (read-once data: do not pollute L1) -------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ int4 ldg_stream_v4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int ldg_stream_s32(const void* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_v4(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream_v2f64(void* p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// ---- warp primitives -----------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(FULL_MASK, v, d);
    if (lane_id() >= unsigned(d)) v += o;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
  return v;
}

// ---- decoupled look-back (single-pass chained scan) ----------------------------------------------
// tile_state[t] packs {flag (2 bits), value (62 bits)} in ONE 64-bit word, so a relaxed 64-bit load
// observes flag and value together and no fence is needed between them.
constexpr unsigned long long LB_INVALID = 0ull;
constexpr unsigned long long LB_PARTIAL = 1ull << 62;   // value = this tile's own aggregate
constexpr unsigned long long LB_PREFIX = 2ull << 62;    // value = inclusive prefix up to this tile
constexpr unsigned long long LB_FLAG_MASK = 3ull << 62;
constexpr unsigned long long LB_VALUE_MASK = ~LB_FLAG_MASK;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ONE full warp of the CTA that owns tile `tile` (tile > 0) after it has published
// (LB_PARTIAL | aggregate).  Returns the exclusive prefix (sum of aggregates of tiles < tile) in every
// lane.  Tiles are handed out by an atomic ticket, so every predecessor has been started by a resident
// CTA and the spin terminates.
__device__ __forceinline__ unsigned long long lookback_exclusive_prefix(const unsigned long long* tile_state, long long tile) {
  unsigned long long exclusive = 0;
  long long base = tile - 1;  // lane L inspects tile base - L
  while (true) {
    long long t = base - (long long)lane_id();
    unsigned long long s;
    if (t >= 0) {
      do {
        s = ld_relaxed_u64(tile_state + t);
      } while ((s & LB_FLAG_MASK) == LB_INVALID);
    } else {
      s = LB_PREFIX;  // virtual tile before the first one: inclusive prefix 0
    }
    unsigned has_prefix = __ballot_sync(FULL_MASK, (s & LB_FLAG_MASK) == LB_PREFIX);
    // lanes nearer than the first PREFIX lane contribute their partials; that lane contributes its prefix
    int first = has_prefix ? __ffs(has_prefix) - 1 : 32;
    unsigned long long contrib = (int(lane_id()) <= first) ? (s & LB_VALUE_MASK) : 0ull;
    exclusive += warp_sum(contrib);
    if (has_prefix) break;
    base -= 32;
  }
  return exclusive;
}

// ---- hashing --------------------------------------------------------------------------------------
// Murmur3 64-bit finaliser (fmix64).  Any hash is legal here: hash values are never observable in
// results.
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
// FNV-1a over bytes, folded through fmix64 (Utf8 keys)
__device__ __forceinline__ uint64_t hash_bytes(const uint8_t* p, int n, uint64_t seed) {
  uint64_t h = 0xcbf29ce484222325ull ^ seed;
  for (int i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  return fmix64(h);
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS: UBLKCP) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared; completes on `bar` with `bytes` transaction bytes.  Both addresses
// and the size must be multiples of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace fg
