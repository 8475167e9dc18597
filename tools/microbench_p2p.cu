// microbench_p2p.cu -- what SM-issued stores into a PEER's memory (NVLink 5 / NVSwitch) sustain, by store width and
// by segment length: the numbers that bound partition_scatter_kernel when its destinations are peer windows.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench_p2p tools/microbench_p2p.cu
// Needs two GPUs with peer access; prints GB/s (best of 5) for a 64 MB transfer per variant.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                               \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) {                                                                \
      std::printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      std::exit(1);                                                                         \
    }                                                                                       \
  } while (0)

// Copies n_bytes in segments of seg_bytes: segment s of the source goes to segment perm(s) of the destination
// (perm = multiplicative shuffle: neighbouring segments land far apart, like a tile's destination runs do).
template <typename V>
__global__ void __launch_bounds__(256) copy_segments(const V* __restrict__ src, V* __restrict__ dst, size_t n_bytes, size_t seg_bytes, int scatter) {
  const size_t n_seg = n_bytes / seg_bytes, per_seg = seg_bytes / sizeof(V);
  const int warps = (gridDim.x * blockDim.x) >> 5, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  for (size_t s = warp; s < n_seg; s += warps) {
    const size_t d = scatter ? (s * 40503ull) % n_seg : s;  // 40503 is odd: a permutation when n_seg is a power of two
    const V* from = src + s * per_seg;
    V* to = dst + d * per_seg;
    for (size_t i = lane; i < per_seg; i += 32) to[i] = from[i];
  }
}

template <typename V>
static double run(const void* src, void* dst, size_t n_bytes, size_t seg, int scatter, int ctas_per_sm, int sms) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(a));
    copy_segments<V><<<sms * ctas_per_sm, 256>>>(static_cast<const V*>(src), static_cast<V*>(dst), n_bytes, seg, scatter);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep) best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return double(n_bytes) / (best * 1e-3) / 1e9;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) {
    std::printf("needs two GPUs\n");
    return 0;
  }
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  std::printf("peer access 0 -> 1: %d\n", can);
  if (!can) return 0;
  const size_t bytes = size_t(64) << 20;
  void *src, *local, *remote;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&remote, bytes));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&src, bytes));
  CK(cudaMalloc(&local, bytes));
  CK(cudaMemset(src, 1, bytes));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  std::printf("%-10s %-8s %-8s %-6s %10s %10s\n", "store", "segment", "scatter", "CTA/SM", "local GB/s", "peer GB/s");
  for (int ctas : {1, 3, 8})
    for (size_t seg : {size_t(256), size_t(1024), size_t(4096), size_t(65536)})
      for (int scatter : {0, 1}) {
        if (!scatter && seg != 65536) continue;
        std::printf("%-10s %-8zu %-8d %-6d %10.0f %10.0f\n", "4 B/lane", seg, scatter, ctas, run<unsigned>(src, local, bytes, seg, scatter, ctas, sms),
                    run<unsigned>(src, remote, bytes, seg, scatter, ctas, sms));
        std::printf("%-10s %-8zu %-8d %-6d %10.0f %10.0f\n", "16 B/lane", seg, scatter, ctas, run<uint4>(src, local, bytes, seg, scatter, ctas, sms),
                    run<uint4>(src, remote, bytes, seg, scatter, ctas, sms));
      }
  // the DMA engine for comparison
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a));
    CK(cudaMemcpyPeerAsync(remote, 1, src, 0, bytes));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = std::min(best, ms);
  }
  std::printf("cudaMemcpyPeerAsync 64 MB: %.0f GB/s\n", double(bytes) / (best * 1e-3) / 1e9);
  return 0;
}
