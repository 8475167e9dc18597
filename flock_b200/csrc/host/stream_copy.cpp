// stream_copy.cpp -- the host copy of the pageable feed: Arrow buffers -> the page-locked staging ring.
//
// A staged feed writes every byte of a relation once into memory the CPU never reads again (the DMA engine does).
// memcpy's ordinary stores first read each destination line into the cache (read-for-ownership) and write it back
// later: three memory transfers per byte moved, and the relation evicts the caches on its way.  Non-temporal stores
// write whole lines past the cache: two transfers.  glibc switches to them only for copies of several megabytes; the
// pieces here are single batches (256 KB).  Plain g++ (no CUDA in this file): built by flock_b200/build.py.
#include <immintrin.h>

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace fg {

namespace {

__attribute__((target("avx2"))) void stream_copy_avx2(char* d, const char* s, size_t n) {
  size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
  if (head > n) head = n;
  std::memcpy(d, s, head);
  d += head;
  s += head;
  n -= head;
  const size_t body = n & ~size_t(127);
  for (size_t i = 0; i < body; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
    const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
  }
  std::memcpy(d + body, s + body, n - body);
  _mm_sfence();  // the streamed lines are globally visible before the caller publishes "this piece is staged"
}

}  // namespace

// Copies n bytes to a destination that will next be read by a device, not by this CPU.  `streaming` = 0 forces memcpy
// (flockgpu_set_option("feed_stream_stores", 0): the A/B switch).
void stage_copy(void* dst, const void* src, size_t n, int streaming) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (streaming && avx2 && n >= 4096) stream_copy_avx2(static_cast<char*>(dst), static_cast<const char*>(src), n);
  else std::memcpy(dst, src, n);
}

}  // namespace fg
