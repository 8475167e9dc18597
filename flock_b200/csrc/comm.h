// comm.h -- the communicator a context carries (comm.cu: NCCL binding + fallback all-to-all; exchange.cu: peer windows).
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "internal.h"

namespace fg {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl();  // comm.cu; fails with FLOCKGPU_ERR_NCCL when the library cannot be loaded

#define FG_NCCL(expr)                                                                                              \
  do {                                                                                                             \
    ncclResult_t _r = (expr);                                                                                      \
    if (_r != ncclSuccess)                                                                                         \
      ::fg::fail(FLOCKGPU_ERR_NCCL, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                                  \
                 ::fg::nccl().GetErrorString ? ::fg::nccl().GetErrorString(_r) : "nccl error");                    \
  } while (0)

constexpr int EX_MAX_WORLD = 16;

// Peer-memory side of the communicator (exchange.cu): every rank owns a receive WINDOW and a small control block,
// both plain cudaMalloc allocations shared with the other ranks of the box through CUDA IPC, so that a partition
// kernel can store rows straight into the receiver's HBM over NVLink.
struct PeerWindows {
  bool enabled = false;
  std::string why_not = "not initialised";
  size_t window_bytes = 0;
  char* win[EX_MAX_WORLD] = {};                 // win[r]: rank r's window in MY address space (own: local pointer)
  unsigned long long* ctrl[EX_MAX_WORLD] = {};  // the same for the control blocks
  unsigned long long* d_result = nullptr;       // device: [0] error, [1] rows received, [2..] bytes received per Utf8 column
  unsigned long long* h_result = nullptr;       // the same, page-locked host copy written by the finish kernel
  unsigned long long seq = 0;                   // exchanges started on this communicator (every rank counts alike)
  // window allocator: regions are handed out front to back and the cursor returns to 0 when no table lives in the
  // window any more
  size_t cursor = 0;
  int64_t live = 0;
};

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  int device = 0;
  PeerWindows pw;
  ~Comm();
};

void peer_windows_init(const CtxPtr& ctx, Comm& cm);  // exchange.cu; never throws: failure leaves pw.enabled false
void peer_windows_destroy(Comm& cm);

}  // namespace fg
