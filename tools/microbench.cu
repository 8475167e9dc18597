// microbench.cu -- B200 (sm_100a) micro-measurements that size the kernels of this repo.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Prints one line per experiment; run under gpurun and keep the output in profiles/.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

__device__ __forceinline__ int4 ldg_stream(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// (1) streaming read, UNROLL independent 16-byte loads per thread in flight
template <int UNROLL>
__global__ void read_ldg(const int4* __restrict__ in, size_t n16, unsigned* sink) {
  unsigned acc = 0;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
    int4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(in + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n16; i += stride) acc += ldg_stream(in + i).x;
  if (acc == 0x12345678u) *sink = acc;
}

// (1b) streaming read through TMA bulk copies (cp.async.bulk -> smem, mbarrier), STAGES x CHUNK bytes in flight per CTA
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES, int CHUNK>
__global__ void read_tma(const char* __restrict__ in, size_t bytes, unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[STAGES];
  const size_t n_chunks = bytes / CHUNK;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned acc = 0;
  size_t first = blockIdx.x, step = gridDim.x;
  // prologue
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      size_t c = first + size_t(s) * step;
      if (c < n_chunks) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + s * CHUNK)),
                     "l"(in + c * CHUNK), "r"(CHUNK), "r"(smem_u32(&bar[s]))
                     : "memory");
      }
    }
  }
  int stage = 0;
  unsigned phase = 0;
  for (size_t c = first; c < n_chunks; c += step) {
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar[stage])),
        "r"(phase)
        : "memory");
    const int4* s4 = reinterpret_cast<const int4*>(smem + stage * CHUNK);
    for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) {
      int4 v = s4[i];
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    __syncthreads();
    size_t nxt = c + size_t(STAGES) * step;
    if (threadIdx.x == 0 && nxt < n_chunks) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[stage])), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + stage * CHUNK)),
                   "l"(in + nxt * CHUNK), "r"(CHUNK), "r"(smem_u32(&bar[stage]))
                   : "memory");
    }
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// (2) shared-memory atomics: MODE 0 distinct addresses per lane, 1 all lanes same address, 2 half of the lanes one hot address
template <int MODE>
__global__ void smem_atomics(unsigned* out, int iters) {
  __shared__ unsigned tab[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = 0;
  __syncthreads();
  unsigned x = threadIdx.x * 2654435761u + blockIdx.x;
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    unsigned idx = MODE == 0 ? ((threadIdx.x & 31) + 32 * ((x >> 20) & 127)) : MODE == 1 ? ((x >> 27) * 0 + 7) : ((x >> 31) ? 7u : ((x >> 12) & 4095));
    atomicAdd(&tab[idx & 4095], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = tab[7];
}

// (3) __match_any_sync throughput
__global__ void match_any(unsigned* out, int iters) {
  unsigned x = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    acc += __match_any_sync(0xffffffffu, (x >> 31) ? 7u : (x >> 25));
  }
  if (acc == 0x12345u) out[0] = acc;
}

// (4) global atomics into a table of `slots` 8-byte counters, random addresses
__global__ void gmem_atomics(unsigned long long* tab, size_t mask, int iters) {
  unsigned long long x = (blockIdx.x * 1024ull + threadIdx.x) * 0x9e3779b97f4a7c15ull + 1;
  for (int it = 0; it < iters; ++it) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33;
    atomicAdd(&tab[x & mask], 1ull);
  }
}


// (5) launch overhead and the skeleton of the single-wave filter: how much of a ~30 us kernel is not the kernel?
struct BigParams {
  unsigned long long words[470];  // ~3.7 KB, the size of FilterArgs
};
__global__ void empty_small(unsigned* sink) {
  if (threadIdx.x == 1025) *sink = 1;
}
__global__ void empty_big(const __grid_constant__ BigParams p, unsigned* sink) {
  if (threadIdx.x == 1025) *sink = unsigned(p.words[blockIdx.x % 470]);
}
// one 16 Ki-row tile per CTA: 16 x 16-byte loads per thread in four rounds, count multiples of 123, then
// (BARRIER) arrival counter + prefix over the counts of all predecessors, as compact.cuh does it
template <bool BARRIER>
__global__ void __launch_bounds__(256, 5) skeleton(const int4* __restrict__ in, size_t n16, unsigned* counts, unsigned* arrived, unsigned target,
                                                   unsigned long long* out) {
  __shared__ unsigned long long wsum[8];
  const int tid = threadIdx.x;
  const size_t base = size_t(blockIdx.x) * 4096;
  unsigned cnt = 0;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    int4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      size_t i = base + size_t(c * 4 + j) * 256 + tid;
      v[j] = i < n16 ? ldg_stream(in + i) : make_int4(1, 1, 1, 1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) cnt += (v[j].x % 123 == 0) + (v[j].y % 123 == 0) + (v[j].z % 123 == 0) + (v[j].w % 123 == 0);
  }
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((tid & 31) == 0) wsum[tid >> 5] = cnt;
  __syncthreads();
  unsigned total = 0;
  for (int w = 0; w < 8; ++w) total += unsigned(wsum[w]);
  if (!BARRIER) {
    if (tid == 0) counts[blockIdx.x] = total;
    return;
  }
  if (tid == 0) {
    counts[blockIdx.x] = total;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(arrived), "r"(1u) : "memory");
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrived) : "memory");
    } while (int(seen - target) < 0);
  }
  __syncthreads();
  unsigned long long part = 0;
  for (unsigned i = tid; i < blockIdx.x; i += 256) part += __ldcg(counts + i);
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  __syncthreads();
  if ((tid & 31) == 0) wsum[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    unsigned long long excl = 0;
    for (int w = 0; w < 8; ++w) excl += wsum[w];
    if (blockIdx.x == gridDim.x - 1) *out = excl + total;
  }
}

template <class F>
float time_ms(F&& f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  return best;
}

int main(int argc, char** argv) {
  const bool only_pcie = argc > 1 && !strcmp(argv[1], "pcie");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  unsigned* sink;
  CK(cudaMalloc(&sink, 4096 * 4));
  const size_t big = size_t(1) << 30;
  char* buf;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 1, big));
  char* flush;
  CK(cudaMalloc(&flush, size_t(512) << 20));

  {
    // (6) zero-copy reads of page-locked HOST memory over PCIe: SM loads (the filter's zero-copy feed) vs TMA bulk
    // copies into shared memory vs one cudaMemcpyAsync of the same 40 MB
    const size_t hbytes = size_t(40) << 20;
    char* hbuf;
    CK(cudaHostAlloc(&hbuf, hbytes, cudaHostAllocDefault));
    memset(hbuf, 1, hbytes);
    auto timed1 = [&](auto&& launch) {
      float best = 1e30f;
      for (int r = 0; r < 4; ++r) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
      }
      return best;
    };
    for (int cta_per_sm : {2, 4, 8}) {
      float t4 = timed1([&] { read_ldg<4><<<sms * cta_per_sm, 256>>>(reinterpret_cast<const int4*>(hbuf), hbytes / 16, sink); });
      float t8 = timed1([&] { read_ldg<8><<<sms * cta_per_sm, 256>>>(reinterpret_cast<const int4*>(hbuf), hbytes / 16, sink); });
      printf("pcie read_ldg 40MB pinned host ctas/sm=%d: unroll4 %.3f ms %.1f GB/s, unroll8 %.3f ms %.1f GB/s\n", cta_per_sm, t4, hbytes / t4 / 1e6, t8, hbytes / t8 / 1e6);
    }
    {
      constexpr int STAGES = 4, CHUNK = 16384;
      CK(cudaFuncSetAttribute(read_tma<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK));
      for (int cta_per_sm : {1, 2, 3}) {
        float t = timed1([&] { read_tma<STAGES, CHUNK><<<sms * cta_per_sm, 256, STAGES * CHUNK>>>(hbuf, hbytes, sink); });
        printf("pcie read_tma 4x16KB 40MB pinned host ctas/sm=%d: %.3f ms %.1f GB/s\n", cta_per_sm, t, hbytes / t / 1e6);
      }
    }
    float tc = timed1([&] { cudaMemcpyAsync(buf, hbuf, hbytes, cudaMemcpyHostToDevice); });
    printf("pcie cudaMemcpyAsync 40MB pinned host -> HBM (one call): %.3f ms %.1f GB/s\n", tc, hbytes / tc / 1e6);
    float tc153 = timed1([&] {
      for (int b = 0; b < 153; ++b) cudaMemcpyAsync(buf + size_t(b) * 262144, hbuf + size_t(b) * 262144, 262144, cudaMemcpyHostToDevice);
    });
    printf("pcie 153 x cudaMemcpyAsync of 256 KB (the per-batch feed): %.3f ms %.1f GB/s\n", tc153, 153 * 262144.0 / tc153 / 1e6);
    CK(cudaGetLastError());
    cudaFreeHost(hbuf);
    if (only_pcie) return 0;
  }

  for (size_t bytes : {size_t(40) << 20, size_t(400) << 20, big}) {
    for (int cta_per_sm : {2, 4, 8}) {
      auto run = [&](auto kernel, const char* name) {
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
          cudaMemsetAsync(flush, r, size_t(512) << 20);  // evict L2
          cudaEvent_t a, b;
          cudaEventCreate(&a);
          cudaEventCreate(&b);
          cudaEventRecord(a);
          kernel<<<sms * cta_per_sm, 256>>>(reinterpret_cast<const int4*>(buf), bytes / 16, sink);
          cudaEventRecord(b);
          cudaEventSynchronize(b);
          float ms;
          cudaEventElapsedTime(&ms, a, b);
          best = ms < best ? ms : best;
        }
        printf("read_ldg %-8s bytes=%4zuMB ctas/sm=%d  %.4f ms  %.0f GB/s\n", name, bytes >> 20, cta_per_sm, best, bytes / best / 1e6);
      };
      run(read_ldg<1>, "unroll1");
      run(read_ldg<4>, "unroll4");
      run(read_ldg<8>, "unroll8");
    }
    {
      constexpr int STAGES = 4, CHUNK = 16384;
      CK(cudaFuncSetAttribute(read_tma<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK));
      for (int cta_per_sm : {1, 2, 3}) {
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
          cudaMemsetAsync(flush, r, size_t(512) << 20);
          cudaEvent_t a, b;
          cudaEventCreate(&a);
          cudaEventCreate(&b);
          cudaEventRecord(a);
          read_tma<STAGES, CHUNK><<<sms * cta_per_sm, 256, STAGES * CHUNK>>>(buf, bytes, sink);
          cudaEventRecord(b);
          cudaEventSynchronize(b);
          float ms;
          cudaEventElapsedTime(&ms, a, b);
          best = ms < best ? ms : best;
        }
        printf("read_tma 4x16KB   bytes=%4zuMB ctas/sm=%d  %.4f ms  %.0f GB/s\n", bytes >> 20, cta_per_sm, best, bytes / best / 1e6);
      }
    }
  }
  CK(cudaGetLastError());


  {
    // launch overhead (event-timed, best of 5, after an L2-evicting memset like the read tests)
    auto timed = [&](auto&& launch) {
      float best = 1e30f;
      for (int r = 0; r < 7; ++r) {
        cudaMemsetAsync(flush, r, size_t(512) << 20);
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
      }
      return best;
    };
    BigParams bp{};
    for (int grid : {148, 611, 740}) {
      float t_small = timed([&] { empty_small<<<grid, 256>>>(sink); });
      float t_big = timed([&] { empty_big<<<grid, 256>>>(bp, sink); });
      void* args_small[] = {&sink};
      float t_coop = timed([&] { cudaLaunchCooperativeKernel(reinterpret_cast<void*>(empty_small), dim3(grid), dim3(256), args_small, 0, nullptr); });
      void* args_big[] = {&bp, &sink};
      float t_coop_big = timed([&] { cudaLaunchCooperativeKernel(reinterpret_cast<void*>(empty_big), dim3(grid), dim3(256), args_big, 0, nullptr); });
      printf("empty kernel grid=%3d x 256: plain %.2f us, 3.7KB params %.2f us, cooperative %.2f us, cooperative+3.7KB %.2f us\n", grid, t_small * 1e3,
             t_big * 1e3, t_coop * 1e3, t_coop_big * 1e3);
    }
    unsigned* counts;
    unsigned* arrived;
    unsigned long long* total;
    CK(cudaMalloc(&counts, 4096 * 4));
    CK(cudaMalloc(&arrived, 64));
    CK(cudaMalloc(&total, 64));
    CK(cudaMemset(arrived, 0, 64));
    const size_t n16 = size_t(10'000'000) / 4;  // 10 M int32 rows = 40 MB
    const int grid = int((n16 + 4095) / 4096);
    unsigned target = 0;
    float t_nb = timed([&] { skeleton<false><<<grid, 256>>>(reinterpret_cast<const int4*>(buf), n16, counts, arrived, 0, total); });
    float t_b = timed([&] {
      target += unsigned(grid);
      skeleton<true><<<grid, 256>>>(reinterpret_cast<const int4*>(buf), n16, counts, arrived, target, total);
    });
    float t_bc = timed([&] {
      target += unsigned(grid);
      const int4* in = reinterpret_cast<const int4*>(buf);
      size_t n = n16;
      void* a[] = {&in, &n, &counts, &arrived, &target, &total};
      cudaLaunchCooperativeKernel(reinterpret_cast<void*>(skeleton<true>), dim3(grid), dim3(256), a, 0, nullptr);
    });
    printf("filter skeleton 10 M rows, grid %d x 256 (one wave): count only %.2f us, + arrival barrier and prefix %.2f us, cooperative %.2f us\n", grid,
           t_nb * 1e3, t_b * 1e3, t_bc * 1e3);
    CK(cudaGetLastError());
  }

  const int iters = 4096;
  {
    float ms0 = time_ms([&] { smem_atomics<0><<<sms * 2, 1024>>>(sink, iters); });
    float ms1 = time_ms([&] { smem_atomics<1><<<sms * 2, 1024>>>(sink, iters); });
    float ms2 = time_ms([&] { smem_atomics<2><<<sms * 2, 1024>>>(sink, iters); });
    double ops = double(sms) * 2 * 1024 * iters;
    printf("smem atomicAdd distinct-per-lane: %.3f ms  %.1f Gop/s (%.2f ops/clk/SM @1.9GHz)\n", ms0, ops / ms0 / 1e6, ops / ms0 / 1e6 / sms / 1.9);
    printf("smem atomicAdd same address     : %.3f ms  %.1f Gop/s\n", ms1, ops / ms1 / 1e6);
    printf("smem atomicAdd 50%% hot key      : %.3f ms  %.1f Gop/s\n", ms2, ops / ms2 / 1e6);
    float ms3 = time_ms([&] { match_any<<<sms * 2, 1024>>>(sink, iters); });
    printf("__match_any_sync                : %.3f ms  %.1f G lane-ops/s\n", ms3, ops / ms3 / 1e6);
  }
  for (size_t slots : {size_t(1) << 16, size_t(1) << 22, size_t(1) << 26}) {
    unsigned long long* tab;
    CK(cudaMalloc(&tab, slots * 8));
    CK(cudaMemset(tab, 0, slots * 8));
    float ms = time_ms([&] { gmem_atomics<<<sms * 4, 512>>>(tab, slots - 1, 256); });
    double ops = double(sms) * 4 * 512 * 256;
    printf("global atomicAdd u64, %8zu slots (%4zu MB): %.3f ms  %.1f Gop/s\n", slots, slots * 8 >> 20, ms, ops / ms / 1e6);
    cudaFree(tab);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
