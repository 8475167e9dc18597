// stage_copy_test.cpp -- the host copy of the pageable feed (flock_b200/csrc/host/stream_copy.cpp), checked on the CPU:
// every combination of source / destination misalignment and sizes around the streaming threshold, no byte written
// outside [dst, dst + n).  Built and run by tests/test_host.py::test_stage_copy_is_exact.
#include <cstdio>
#include <cstring>
#include <vector>
#include <cstdlib>
namespace fg { void stage_copy(void* dst, const void* src, size_t n, int streaming); }
int main() {
  const size_t N = 64u << 20;
  std::vector<char> a(N + 64), b(N + 128), c(N + 128);
  for (size_t i = 0; i < a.size(); ++i) a[i] = char(i * 2654435761u >> 13);
  // correctness at odd alignments / sizes
  for (size_t off_s : {0, 1, 7, 33}) for (size_t off_d : {0, 3, 31, 64}) for (size_t n : {size_t(0), size_t(1), size_t(4095), size_t(4096), size_t(4097), size_t(262144), size_t(1000003)}) {
    memset(b.data(), 0x55, n + off_d + 64);
    fg::stage_copy(b.data() + off_d, a.data() + off_s, n, 1);
    if (memcmp(b.data() + off_d, a.data() + off_s, n) != 0) { printf("MISMATCH %zu %zu %zu\n", off_s, off_d, n); return 1; }
    for (size_t k = 0; k < 32; ++k) if (b[off_d + n + k] != 0x55) { printf("OVERRUN %zu %zu %zu\n", off_s, off_d, n); return 1; }
    for (size_t k = 0; k < off_d; ++k) if (b[k] != 0x55) { printf("UNDERRUN\n"); return 1; }
  }
  printf("ok\n");
}
