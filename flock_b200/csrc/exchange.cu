// exchange.cu -- RepartitionExec(Hash(keys, world)) + the shuffle between stages as ONE fused partition-and-push over
// NVLink peer memory.
//
// In the reference the shuffle is N x M AWS Lambda invocations carrying Arrow-Flight payloads
// (flock-function/src/aws/actor.rs:425-543: for every output partition of the stage, `invoke_next_functions` ships that
// partition to the function that owns it; plan shape flock/src/distributed_plan/planner.rs:152-171, stage.rs:597-601).
// Round 1 did it with one compaction pass and one take() per destination, a size all-gather with a host round trip and
// a grouped NCCL send/recv: 165 kernel launches on rank 0 for q8 at 8 GPUs, 20 % weak-scaling efficiency.
//
// Now every rank owns a receive WINDOW (plain cudaMalloc memory shared with the other ranks of the box through CUDA
// IPC) and an exchange is five launches, none of them per destination, with one host wait at the end:
//   partition_count_kernel    destinations + per-CTA histograms                          (partition.cu)
//   partition_scan_kernel     scans over the CTAs, my row / byte counts per destination   (partition.cu)
//   exchange_place_kernel     the count all-gather WITHOUT the host: every rank stores its counts (and where in its
//                             window the next exchange may land) into every peer's mailbox over NVLink, waits for the
//                             W rows of the matrix, and derives from it -- identically on every rank -- the layout of
//                             every receiver's region and where its own rows go there
//   partition_scatter_kernel  orders 2048 rows at a time by destination in shared memory and stores them straight into
//                             the receivers' windows: the transfer IS the partition kernel's store phase
//   exchange_finish_kernel    tells every receiver "my rows are in" and waits for the same word from every source
// The received relation is a set of VIEWS into the window (no copy out): the region is released when the last table
// that lives in it dies.  No data-path NCCL call remains; NCCL still bootstraps the communicator (and carries the
// fallback all_to_all() when peer windows cannot be set up, e.g. two contexts inside one process).
//
// Ordering argument.  (1) A receiver's window region is reused only after the receiver itself has placed a later
// exchange there, and a sender writes into a region only after it has seen the receiver's mailbox row for THAT
// exchange, which the receiver's place kernel wrote in stream order after every consumer of the previous tenant.
// (2) Mailbox rows are double-buffered by exchange parity: a rank can be at most one exchange ahead of the slowest
// reader of its row, because finishing an exchange needs every rank's row of that exchange.  (3) Data stores are
// ordered before the "rows are in" word by the kernel boundary plus a system-scope fence; the receiver acquires it.
#include <algorithm>

#include "comm.h"
#include "device_utils.cuh"
#include "partition.h"

namespace fg {

// ---- control block (u64 words) ---------------------------------------------------------------------------------
constexpr int EX_MSG_HEAD = 4;                                                    // [0] tag = exchange number, [1] region offset, [2] region capacity
constexpr int EX_MSG_WORDS = EX_MSG_HEAD + EX_MAX_WORLD * (1 + PT_MAX_UTF8);      // then rows[d], bytes[u][d]
constexpr int EX_MAILBOX_WORDS = 2 * EX_MAX_WORLD * EX_MSG_WORDS;                 // [parity][source]
constexpr int EX_DONE_AT = EX_MAILBOX_WORDS;                                      // done[source] = last exchange whose rows are in
constexpr int EX_CTRL_WORDS = EX_DONE_AT + EX_MAX_WORLD;
constexpr unsigned long long EX_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;         // a peer that never shows up: fail, do not hang

enum ExError : unsigned long long { EX_OK = 0, EX_ERR_TIMEOUT = 1, EX_ERR_WINDOW = 2 };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__host__ __device__ __forceinline__ unsigned long long align256(unsigned long long x) { return (x + 255ull) & ~255ull; }

// Layout of a receiver's region, a pure function of what it receives in total: every rank computes it alike.
struct ExLayout {
  unsigned long long val[MAX_IN_COLS];
  unsigned long long off[PT_MAX_UTF8];
  unsigned long long bytes[PT_MAX_UTF8];
  unsigned long long end;
};
__host__ __device__ inline ExLayout exchange_layout(int n_fixed, const int32_t* fwidth, int n_utf8, unsigned long long rows, const unsigned long long* bytes) {
  ExLayout l{};
  unsigned long long o = 0;
  for (int f = 0; f < n_fixed; ++f) {
    l.val[f] = o;
    o += align256(rows * (unsigned long long)fwidth[f] + 16);  // + 16: vector loads may overrun the tail
  }
  for (int u = 0; u < n_utf8; ++u) {
    l.off[u] = o;
    o += align256((rows + 1) * 4 + 16);
    l.bytes[u] = o;
    o += align256(bytes[u] + 16);
  }
  l.end = o;
  return l;
}

struct ExPlaceArgs {
  unsigned long long seq;
  int32_t me, world, n_fixed, n_utf8;
  int32_t fwidth[MAX_IN_COLS];
  const unsigned long long* my_counts;  // [(1 + n_utf8)][world] from partition_scan_kernel
  unsigned long long region_off, region_cap;
  unsigned long long* ctrl[EX_MAX_WORLD];
  char* win[EX_MAX_WORLD];
  PartDest* dest;
  unsigned long long* result;  // [0] error, [1] rows I receive, [2 + u] bytes I receive of Utf8 column u
};

__global__ void __launch_bounds__(64) exchange_place_kernel(const __grid_constant__ ExPlaceArgs a) {
  __shared__ unsigned s_err;
  const int t = threadIdx.x, W = a.world, U = a.n_utf8;
  if (t == 0) s_err = EX_OK;
  __syncthreads();
  const int parity = int(a.seq & 1);
  // ---- my row of the matrix goes to every rank (myself included), payload first, tag last
  if (t < W) {
    unsigned long long* box = a.ctrl[t] + (size_t(parity) * EX_MAX_WORLD + a.me) * EX_MSG_WORDS;
    box[1] = a.region_off;
    box[2] = a.region_cap;
    for (int k = 0; k < 1 + U; ++k)
      for (int d = 0; d < W; ++d) box[EX_MSG_HEAD + k * EX_MAX_WORLD + d] = a.my_counts[k * W + d];
    __threadfence_system();
    st_release_sys(box, a.seq);
  }
  // ---- wait for every rank's row in MY mailbox
  const unsigned long long* mail = a.ctrl[a.me] + size_t(parity) * EX_MAX_WORLD * EX_MSG_WORDS;
  if (t < W) {
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(mail + size_t(t) * EX_MSG_WORDS) != a.seq) {
      if (global_timer_ns() - t0 > EX_TIMEOUT_NS) {
        atomicMax(&s_err, unsigned(EX_ERR_TIMEOUT));
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (s_err != EX_OK) {
    if (t == 0) a.result[0] = s_err;
    return;
  }
  // ---- thread d: what receiver d gets in total, where its region is, where MY rows go there
  if (t < W) {
    const int d = t;
    unsigned long long rows_before = 0, rows_total = 0, bytes_before[PT_MAX_UTF8] = {}, bytes_total[PT_MAX_UTF8] = {};
    for (int s = 0; s < W; ++s) {
      const unsigned long long* row = mail + size_t(s) * EX_MSG_WORDS + EX_MSG_HEAD;
      const unsigned long long r = row[d];
      if (s < a.me) rows_before += r;
      rows_total += r;
      for (int u = 0; u < U; ++u) {
        const unsigned long long b = row[(1 + u) * EX_MAX_WORLD + d];
        if (s < a.me) bytes_before[u] += b;
        bytes_total[u] += b;
      }
    }
    const unsigned long long* drow = mail + size_t(d) * EX_MSG_WORDS;
    const unsigned long long region = drow[1], cap = drow[2];
    const ExLayout lay = exchange_layout(a.n_fixed, a.fwidth, U, rows_total, bytes_total);
    bool fits = lay.end <= cap && rows_total < (1ull << 32) - 1;
    for (int u = 0; u < U; ++u) fits = fits && bytes_total[u] < (1ull << 31);
    if (!fits) {
      atomicMax(&s_err, unsigned(EX_ERR_WINDOW));
    } else {
      char* base = a.win[d] + region;
      PartDest pd{};
      for (int f = 0; f < a.n_fixed; ++f) pd.val[f] = base + lay.val[f] + rows_before * (unsigned long long)a.fwidth[f];
      for (int u = 0; u < U; ++u) {
        pd.off[u] = reinterpret_cast<int32_t*>(base + lay.off[u]) + rows_before;
        pd.bytes[u] = reinterpret_cast<uint8_t*>(base + lay.bytes[u]) + bytes_before[u];
        pd.byte_origin[u] = (long long)bytes_before[u];
        // the terminal offsets entry of the received column is written by the last source
        if (a.me == W - 1) reinterpret_cast<int32_t*>(base + lay.off[u])[rows_total] = int32_t(bytes_total[u]);
      }
      a.dest[d] = pd;
      if (d == a.me) {
        a.result[1] = rows_total;
        for (int u = 0; u < U; ++u) a.result[2 + u] = bytes_total[u];
      }
    }
  }
  __syncthreads();
  if (t == 0) a.result[0] = s_err;  // partition_scatter_kernel returns at once when this is non-zero
}

struct ExFinishArgs {
  unsigned long long seq;
  int32_t me, world, n_words, pad;
  unsigned long long* ctrl[EX_MAX_WORLD];
  unsigned long long* result;       // device
  unsigned long long* host_result;  // page-locked copy the host reads after the stream has drained
};

__global__ void __launch_bounds__(64) exchange_finish_kernel(const __grid_constant__ ExFinishArgs a) {
  __shared__ unsigned s_err;
  const int t = threadIdx.x, W = a.world;
  if (t == 0) s_err = EX_OK;
  __syncthreads();
  // the scatter kernel has completed (stream order): its stores are performed; order them before the word below
  if (t < W) {
    __threadfence_system();
    st_release_sys(a.ctrl[t] + EX_DONE_AT + a.me, a.seq);
  }
  if (t < W) {
    const unsigned long long* done = a.ctrl[a.me] + EX_DONE_AT + t;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(done) < a.seq) {
      if (global_timer_ns() - t0 > EX_TIMEOUT_NS) {
        atomicMax(&s_err, unsigned(EX_ERR_TIMEOUT));
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (t == 0) {
    if (a.result[0] == EX_OK && s_err != EX_OK) a.result[0] = s_err;
    for (int i = a.n_words - 1; i >= 0; --i) reinterpret_cast<volatile unsigned long long*>(a.host_result)[i] = a.result[i];
  }
}

// ================================================================================================
// set-up / tear-down of the windows
// ================================================================================================
static size_t window_bytes_wanted(const CtxPtr& ctx) {
  long long mb = ctx->exchange_window_mb;
  if (const char* e = getenv("FLOCKGPU_WINDOW_MB")) mb = atoll(e);
  if (mb < 16) mb = 16;
  return size_t(mb) << 20;
}

void peer_windows_destroy(Comm& cm) {
  PeerWindows& pw = cm.pw;
  for (int r = 0; r < cm.world && r < EX_MAX_WORLD; ++r) {
    if (r == cm.rank) continue;
    if (pw.win[r]) cudaIpcCloseMemHandle(pw.win[r]);
    if (pw.ctrl[r]) cudaIpcCloseMemHandle(pw.ctrl[r]);
    pw.win[r] = nullptr;
    pw.ctrl[r] = nullptr;
  }
  if (cm.rank < EX_MAX_WORLD) {
    if (pw.win[cm.rank]) cudaFree(pw.win[cm.rank]);
    if (pw.ctrl[cm.rank]) cudaFree(pw.ctrl[cm.rank]);
    pw.win[cm.rank] = nullptr;
    pw.ctrl[cm.rank] = nullptr;
  }
  if (pw.d_result) cudaFree(pw.d_result);
  if (pw.h_result) cudaFreeHost(pw.h_result);
  pw.d_result = pw.h_result = nullptr;
  pw.enabled = false;
}

void peer_windows_init(const CtxPtr& ctx, Comm& cm) {
  PeerWindows& pw = cm.pw;
  const int W = cm.world, me = cm.rank;
  struct Card {
    cudaIpcMemHandle_t win, ctrl;
    int ok;
    int pad[3];
  };
  Card mine{};
  std::string why;
  const char* mode = getenv("FLOCKGPU_EXCHANGE");
  if (W < 2) why = "single rank";
  else if (W > EX_MAX_WORLD) why = "more than 16 ranks";
  else if (mode && !strcmp(mode, "nccl")) why = "FLOCKGPU_EXCHANGE=nccl";
  if (why.empty()) {
    pw.window_bytes = window_bytes_wanted(ctx);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&pw.win[me]), pw.window_bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&pw.ctrl[me]), EX_CTRL_WORDS * 8);
    if (e == cudaSuccess) e = cudaMemset(pw.ctrl[me], 0, EX_CTRL_WORDS * 8);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&pw.d_result), 16 * 8);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&pw.h_result), 16 * 8, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.win, pw.win[me]);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.ctrl, pw.ctrl[me]);
    if (e != cudaSuccess) {
      why = std::string("window allocation / IPC export failed: ") + cudaGetErrorString(e);
      cudaGetLastError();
    }
  }
  mine.ok = why.empty() ? 1 : 0;
  // every rank takes part in both all-gathers whatever its own outcome: the decision must be unanimous
  std::vector<Card> all(size_t(std::max(W, 1)));
  bool gathered = false;
  try {
    if (W >= 2) {
      BufferPtr d_mine = alloc(ctx, sizeof(Card)), d_all = alloc(ctx, sizeof(Card) * W);
      FG_CUDA(cudaMemcpyAsync(d_mine->ptr, &mine, sizeof(Card), cudaMemcpyHostToDevice, ctx->stream));
      FG_NCCL(nccl().AllGather(d_mine->ptr, d_all->ptr, sizeof(Card), ncclChar, cm.comm, ctx->stream));
      FG_CUDA(cudaMemcpyAsync(all.data(), d_all->ptr, sizeof(Card) * W, cudaMemcpyDeviceToHost, ctx->stream));
      FG_CUDA(cudaStreamSynchronize(ctx->stream));
      gathered = true;
      int opened = 1;
      for (int r = 0; r < W; ++r) opened &= all[r].ok;
      if (opened) {
        for (int r = 0; r < W && opened; ++r) {
          if (r == me) continue;
          void *w = nullptr, *c = nullptr;
          cudaError_t e = cudaIpcOpenMemHandle(&w, all[r].win, cudaIpcMemLazyEnablePeerAccess);
          if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&c, all[r].ctrl, cudaIpcMemLazyEnablePeerAccess);
          if (e != cudaSuccess) {
            if (why.empty()) why = std::string("cudaIpcOpenMemHandle failed for rank ") + std::to_string(r) + ": " + cudaGetErrorString(e);
            cudaGetLastError();
            opened = 0;
          }
          pw.win[r] = static_cast<char*>(w);
          pw.ctrl[r] = static_cast<unsigned long long*>(c);
        }
      } else if (why.empty()) {
        why = "a peer could not export its window";
      }
      // second round: did everybody manage to map everybody?
      int flag = opened;
      std::vector<int> flags(size_t(W) * 4);
      BufferPtr d_flag = alloc(ctx, 16), d_flags = alloc(ctx, size_t(16) * W);
      int mine4[4] = {flag, 0, 0, 0};
      FG_CUDA(cudaMemcpyAsync(d_flag->ptr, mine4, 16, cudaMemcpyHostToDevice, ctx->stream));
      FG_NCCL(nccl().AllGather(d_flag->ptr, d_flags->ptr, 16, ncclChar, cm.comm, ctx->stream));
      FG_CUDA(cudaMemcpyAsync(flags.data(), d_flags->ptr, size_t(16) * W, cudaMemcpyDeviceToHost, ctx->stream));
      FG_CUDA(cudaStreamSynchronize(ctx->stream));
      for (int r = 0; r < W; ++r)
        if (!flags[size_t(r) * 4] && why.empty()) why = "rank " + std::to_string(r) + " could not map a peer window";
    }
  } catch (const Error& e) {
    if (why.empty()) why = e.msg;
  }
  (void)gathered;
  if (why.empty()) {
    pw.enabled = true;
    pw.why_not.clear();
  } else {
    const size_t wb = pw.window_bytes;
    peer_windows_destroy(cm);
    pw.window_bytes = wb;
    pw.why_not = why;
  }
}

// ================================================================================================
// the exchange
// ================================================================================================
namespace {

// A table that lives in the window holds one lease; the window's cursor returns to 0 with the last one.
struct WindowLease {
  std::shared_ptr<Comm> comm;
  ~WindowLease() {
    PeerWindows& pw = comm->pw;
    if (--pw.live == 0) pw.cursor = 0;
  }
};

TablePtr exchange_fallback(const CtxPtr& ctx, const TablePtr& in, const std::vector<int>& keys, int dest) {
  const int W = ctx->comm->world;
  if (dest >= 0) {
    std::vector<TablePtr> parts;
    for (int r = 0; r < W; ++r) parts.push_back(r == dest ? in : empty_like(ctx, *in));
    return all_to_all(ctx, parts);
  }
  return all_to_all(ctx, hash_partition(ctx, in, keys, W));
}

}  // namespace

TablePtr hash_exchange(const CtxPtr& ctx, const TablePtr& in_ptr, const std::vector<int>& keys, int dest) {
  if (!ctx->comm || ctx->comm->world == 1) return in_ptr;
  Comm& cm = *ctx->comm;
  PeerWindows& pw = cm.pw;
  const int W = cm.world, me = cm.rank;
  const Table& in = *in_ptr;
  in.dense();
  FG_CHECK(dest < W, FLOCKGPU_ERR_INVALID, "hash_exchange: destination %d of %d ranks", dest, W);
  std::vector<int> routing = dest < 0 ? routing_columns(in, keys) : std::vector<int>{};
  std::vector<std::string> routed_on;
  for (int k : routing) routed_on.push_back(in.cols[k].name);
  size_t n_utf8 = 0;
  for (const Column& c : in.cols) n_utf8 += c.dtype == FLOCKGPU_UTF8;
  TablePtr result;
  if (!pw.enabled || n_utf8 > size_t(PT_MAX_UTF8)) {
    result = exchange_fallback(ctx, in_ptr, dest < 0 ? routing : keys, dest);
  } else {
    PartPass ps = partition_count_scan(ctx, in, routing, W, dest, -1, 0, /*ship_nullable=*/true);
    const int U = int(ps.utf8_cols.size());
    const unsigned long long seq = ++pw.seq;
    const size_t region_off = (pw.cursor + 255) & ~size_t(255);
    FG_CHECK(region_off < pw.window_bytes, FLOCKGPU_ERR_UNSUPPORTED,
             "hash_exchange: the %zu MB receive window is full of live tables (raise it with flockgpu_set_option(\"exchange_window_mb\") before comm_init)",
             pw.window_bytes >> 20);
    ExPlaceArgs pa{};
    pa.seq = seq;
    pa.me = me;
    pa.world = W;
    pa.n_fixed = int(ps.fixed_cols.size() + ps.valid_cols.size());  // validity bytes travel as 1-byte columns behind the values
    pa.n_utf8 = U;
    for (size_t f = 0; f < ps.fixed_cols.size(); ++f) pa.fwidth[f] = in.cols[ps.fixed_cols[f]].width();
    for (size_t k = 0; k < ps.valid_cols.size(); ++k) pa.fwidth[ps.fixed_cols.size() + k] = 1;
    pa.my_counts = ps.totals->as<unsigned long long>();
    pa.region_off = region_off;
    pa.region_cap = pw.window_bytes - region_off;
    for (int r = 0; r < W; ++r) {
      pa.ctrl[r] = pw.ctrl[r];
      pa.win[r] = pw.win[r];
    }
    pa.dest = ps.dest->as<PartDest>();
    pa.result = pw.d_result;
    {
      LaunchTimer lt(ctx, "exchange_place_kernel");
      exchange_place_kernel<<<1, 64, 0, ctx->stream>>>(pa);
    }
    FG_CUDA(cudaGetLastError());
    count_launch(ctx);
    partition_scatter(ctx, in, ps, reinterpret_cast<const unsigned*>(pw.d_result));
    ExFinishArgs fa{};
    fa.seq = seq;
    fa.me = me;
    fa.world = W;
    fa.n_words = 2 + U;
    for (int r = 0; r < W; ++r) fa.ctrl[r] = pw.ctrl[r];
    fa.result = pw.d_result;
    fa.host_result = pw.h_result;
    {
      LaunchTimer lt(ctx, "exchange_finish_kernel");
      exchange_finish_kernel<<<1, 64, 0, ctx->stream>>>(fa);
    }
    FG_CUDA(cudaGetLastError());
    count_launch(ctx);
    // the one host wait of an exchange: how much did I receive?
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long err = pw.h_result[0];
    FG_CHECK(err != EX_ERR_TIMEOUT, FLOCKGPU_ERR_NCCL, "hash_exchange %llu: a peer did not show up within %llu s", seq, EX_TIMEOUT_NS / 1000000000ull);
    FG_CHECK(err != EX_ERR_WINDOW, FLOCKGPU_ERR_UNSUPPORTED,
             "hash_exchange %llu: a receiver's %zu MB window cannot hold what it is sent (raise \"exchange_window_mb\" before comm_init)", seq, pw.window_bytes >> 20);
    FG_CHECK(err == EX_OK, FLOCKGPU_ERR_NCCL, "hash_exchange %llu: error %llu", seq, err);
    const unsigned long long rows = pw.h_result[1];
    unsigned long long bytes[PT_MAX_UTF8] = {};
    for (int u = 0; u < U; ++u) bytes[u] = pw.h_result[2 + u];
    const ExLayout lay = exchange_layout(pa.n_fixed, pa.fwidth, U, rows, bytes);
    auto lease = std::make_shared<WindowLease>();
    lease->comm = ctx->comm;
    ++pw.live;
    pw.cursor = region_off + size_t(lay.end);
    std::shared_ptr<const void> owner = std::static_pointer_cast<const void>(lease);
    char* base = pw.win[me] + region_off;
    auto t = std::make_shared<Table>();
    t->ctx = ctx;
    t->metadata = in.metadata;
    t->num_rows = int64_t(rows);
    t->cols.resize(in.cols.size());
    for (size_t f = 0; f < ps.fixed_cols.size(); ++f) {
      const Column& src = in.cols[ps.fixed_cols[f]];
      Column& c = t->cols[ps.fixed_cols[f]];
      c.dtype = src.dtype;
      c.name = src.name;
      c.format = src.format;
      c.nullable = src.nullable;
      c.length = int64_t(rows);
      c.data = std::make_shared<Buffer>(ctx, base + lay.val[f], size_t(rows) * src.width(), owner);
    }
    for (int u = 0; u < U; ++u) {
      const Column& src = in.cols[ps.utf8_cols[u]];
      Column& c = t->cols[ps.utf8_cols[u]];
      c.dtype = src.dtype;
      c.name = src.name;
      c.format = src.format;
      c.nullable = src.nullable;
      c.length = int64_t(rows);
      c.offsets = std::make_shared<Buffer>(ctx, base + lay.off[u], size_t(rows + 1) * 4, owner);
      c.data = std::make_shared<Buffer>(ctx, base + lay.bytes[u], size_t(bytes[u]), owner);
      c.values_bytes = int64_t(bytes[u]);
    }
    for (size_t k = 0; k < ps.valid_cols.size(); ++k)
      t->cols[ps.valid_cols[k]].validity = std::make_shared<Buffer>(ctx, base + lay.val[ps.fixed_cols.size() + k], size_t(rows), owner);
    result = t;
  }
  if (dest < 0) {
    auto r = std::make_shared<Table>(*result);
    r->partitioned_on = routed_on;
    r->partition_world = W;
    result = r;
  }
  return result;
}

}  // namespace fg

using namespace fg;

extern "C" int flockgpu_hash_exchange(flockgpu_ctx* ctx, const flockgpu_table* in, const int32_t* key_cols, int32_t n_keys, flockgpu_table** out) {
  return guarded([&] {
    auto c = core_of(ctx);
    FG_CHECK(out && in && in->table && key_cols && n_keys > 0, FLOCKGPU_ERR_INVALID, "hash_exchange: bad arguments");
    std::lock_guard<std::recursive_mutex> g(c->mu);
    FG_CUDA(cudaSetDevice(c->device));
    std::vector<int> keys(key_cols, key_cols + n_keys);
    for (int k : keys) FG_CHECK(k >= 0 && k < int(in->table->cols.size()), FLOCKGPU_ERR_INVALID, "hash_exchange: key column %d out of range", k);
    *out = wrap_table(hash_exchange(c, in->table, keys));
  });
}
