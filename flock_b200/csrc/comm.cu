// comm.cu -- the hash shuffle between stages as ONE all-to-all over NVLink (NCCL).
//
// In the reference the shuffle is N x M AWS Lambda invocations carrying zstd-compressed Arrow-Flight
// payloads (flock-function/src/aws/actor.rs:425-543; flock/src/aws/lambda.rs:59-128).  Here the hash
// partitions of every rank are exchanged with a single grouped ncclSend/ncclRecv per column buffer:
// sizes first (one small all-gather of the row/byte counts), then the payload.
//
// NCCL is bound with dlopen/dlsym so that libflockgpu.so carries no link-time NCCL dependency: inside
// a torch process the already-loaded (torch-bundled) libnccl.so.2 is reused, stand-alone the system
// library is loaded.
#include <algorithm>

#include "comm.h"

namespace fg {

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define FG_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
    FG_SYM(GetUniqueId, "ncclGetUniqueId");
    FG_SYM(CommInitRank, "ncclCommInitRank");
    FG_SYM(CommDestroy, "ncclCommDestroy");
    FG_SYM(GroupStart, "ncclGroupStart");
    FG_SYM(GroupEnd, "ncclGroupEnd");
    FG_SYM(Send, "ncclSend");
    FG_SYM(Recv, "ncclRecv");
    FG_SYM(AllGather, "ncclAllGather");
    FG_SYM(GetErrorString, "ncclGetErrorString");
#undef FG_SYM
  });
  FG_CHECK(api.handle && api.GetUniqueId && api.CommInitRank && api.Send && api.Recv && api.AllGather && api.GroupStart && api.GroupEnd,
           FLOCKGPU_ERR_NCCL, "NCCL is not available (dlopen libnccl.so.2 failed: %s)", dlerror() ? dlerror() : "missing symbols");
  return api;
}

Comm::~Comm() {
  cudaSetDevice(device);
  peer_windows_destroy(*this);
  if (comm && nccl().CommDestroy) nccl().CommDestroy(comm);
}

void comm_unique_id(uint8_t* out) {
  static_assert(sizeof(ncclUniqueId) <= FLOCKGPU_UNIQUE_ID_BYTES, "ncclUniqueId does not fit the ABI buffer");
  ncclUniqueId id;
  FG_NCCL(nccl().GetUniqueId(&id));
  memset(out, 0, FLOCKGPU_UNIQUE_ID_BYTES);
  memcpy(out, &id, sizeof id);
}

void comm_init(const CtxPtr& ctx, const uint8_t* idbytes, int rank, int world) {
  FG_CHECK(world >= 1 && rank >= 0 && rank < world, FLOCKGPU_ERR_INVALID, "comm_init: bad rank %d / world %d", rank, world);
  ncclUniqueId id;
  memcpy(&id, idbytes, sizeof id);
  auto c = std::make_shared<Comm>();
  c->rank = rank;
  c->world = world;
  c->device = ctx->device;
  FG_NCCL(nccl().CommInitRank(&c->comm, world, id, rank));
  ctx->comm = c;
  peer_windows_init(ctx, *c);  // NVLink peer windows for the exchange; on failure the NCCL all-to-all below stays in charge
}

int comm_world(const CtxPtr& ctx) { return ctx->comm ? ctx->comm->world : 1; }
int comm_rank(const CtxPtr& ctx) { return ctx->comm ? ctx->comm->rank : 0; }

__global__ void add_offset_kernel(int32_t* offs, int64_t n, int32_t delta) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) offs[i] += delta;
}
__global__ void store_i32_kernel(int32_t* p, int32_t v) { *p = v; }

TablePtr all_to_all(const CtxPtr& ctx, const std::vector<TablePtr>& parts) {
  FG_CHECK(ctx->comm, FLOCKGPU_ERR_INVALID, "all_to_all: flockgpu_comm_init has not been called on this context");
  Comm& cm = *ctx->comm;
  const int W = cm.world;
  FG_CHECK(int(parts.size()) == W, FLOCKGPU_ERR_INVALID, "all_to_all: %zu partitions for %d ranks", parts.size(), W);
  for (const TablePtr& p : parts) p->dense();
  const Table& proto = *parts[0];
  const size_t ncol = proto.cols.size();
  std::vector<int> utf8_cols;
  for (size_t c = 0; c < ncol; ++c)
    if (proto.cols[c].dtype == FLOCKGPU_UTF8) utf8_cols.push_back(int(c));
  for (const TablePtr& p : parts) {
    FG_CHECK(p->cols.size() == ncol, FLOCKGPU_ERR_INVALID, "all_to_all: partitions differ in column count");
    for (size_t c = 0; c < ncol; ++c) {
      FG_CHECK(p->cols[c].dtype == proto.cols[c].dtype, FLOCKGPU_ERR_INVALID, "all_to_all: partitions differ in column types");
      FG_CHECK(!p->cols[c].all_null, FLOCKGPU_ERR_UNSUPPORTED, "all_to_all: NULL column");
      require_no_nulls(p->cols[c], "all_to_all (NCCL fallback; the peer-window exchange carries validity)");
    }
  }
  // ---- sizes: M values per destination = rows + value bytes of every Utf8 column
  const int M = 1 + int(utf8_cols.size());
  std::vector<long long> send_meta(size_t(W) * M);
  for (int r = 0; r < W; ++r) {
    send_meta[size_t(r) * M] = parts[r]->num_rows;
    for (size_t u = 0; u < utf8_cols.size(); ++u) send_meta[size_t(r) * M + 1 + u] = parts[r]->cols[utf8_cols[u]].values_bytes;
  }
  BufferPtr d_send = alloc(ctx, send_meta.size() * 8), d_all = alloc(ctx, send_meta.size() * 8 * W);
  FG_CUDA(cudaMemcpyAsync(d_send->ptr, send_meta.data(), send_meta.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  FG_NCCL(nccl().AllGather(d_send->ptr, d_all->ptr, send_meta.size(), ncclInt64, cm.comm, ctx->stream));
  std::vector<long long> all_meta(send_meta.size() * W);
  FG_CUDA(cudaMemcpyAsync(all_meta.data(), d_all->ptr, all_meta.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  FG_CUDA(cudaStreamSynchronize(ctx->stream));
  // all_meta[s][r][m]: what rank s sends to rank r
  auto meta = [&](int s, int r, int m) { return all_meta[(size_t(s) * W + r) * M + m]; };
  std::vector<int64_t> recv_rows(W), row_base(W + 1, 0);
  for (int s = 0; s < W; ++s) {
    recv_rows[s] = meta(s, cm.rank, 0);
    row_base[s + 1] = row_base[s] + recv_rows[s];
  }
  const int64_t total_rows = row_base[W];
  FG_CHECK(total_rows < (int64_t(1) << 32) - 1, FLOCKGPU_ERR_UNSUPPORTED, "all_to_all: more than 2^32-2 rows received");

  auto out = std::make_shared<Table>();
  out->ctx = ctx;
  out->metadata = proto.metadata;
  out->num_rows = total_rows;
  std::vector<std::vector<int64_t>> byte_base(utf8_cols.size(), std::vector<int64_t>(W + 1, 0));
  for (size_t c = 0; c < ncol; ++c) {
    Column col;
    col.dtype = proto.cols[c].dtype;
    col.name = proto.cols[c].name;
    col.format = proto.cols[c].format;
    col.nullable = proto.cols[c].nullable;
    col.length = total_rows;
    if (col.dtype != FLOCKGPU_UTF8) {
      col.data = alloc(ctx, size_t(total_rows) * col.width());
    } else {
      size_t u = std::find(utf8_cols.begin(), utf8_cols.end(), int(c)) - utf8_cols.begin();
      for (int s = 0; s < W; ++s) byte_base[u][s + 1] = byte_base[u][s] + meta(s, cm.rank, 1 + int(u));
      FG_CHECK(byte_base[u][W] < (int64_t(1) << 31), FLOCKGPU_ERR_UNSUPPORTED, "all_to_all: Utf8 column exceeds 2^31-1 bytes");
      col.offsets = alloc(ctx, size_t(total_rows + 1) * 4);
      col.data = alloc(ctx, size_t(byte_base[u][W]));
      col.values_bytes = byte_base[u][W];
    }
    out->cols.push_back(std::move(col));
  }
  // ---- payload: one grouped exchange for every buffer of every column
  FG_NCCL(nccl().GroupStart());
  for (size_t c = 0; c < ncol; ++c) {
    Column& oc = out->cols[c];
    const bool is_utf8 = oc.dtype == FLOCKGPU_UTF8;
    const size_t u = is_utf8 ? size_t(std::find(utf8_cols.begin(), utf8_cols.end(), int(c)) - utf8_cols.begin()) : 0;
    for (int r = 0; r < W; ++r) {
      const Column& sc = parts[r]->cols[c];
      if (!is_utf8) {
        const int w = oc.width();
        if (sc.length) FG_NCCL(nccl().Send(sc.values(), size_t(sc.length) * w, ncclChar, r, cm.comm, ctx->stream));
        if (recv_rows[r])
          FG_NCCL(nccl().Recv(static_cast<char*>(oc.data->ptr) + row_base[r] * w, size_t(recv_rows[r]) * w, ncclChar, r, cm.comm, ctx->stream));
      } else {
        // offsets travel relative to the sender's partition (they start at 0 there); rebased below
        if (sc.length) FG_NCCL(nccl().Send(sc.offs(), size_t(sc.length) * 4, ncclChar, r, cm.comm, ctx->stream));
        if (recv_rows[r]) FG_NCCL(nccl().Recv(oc.offsets->as<int32_t>() + row_base[r], size_t(recv_rows[r]) * 4, ncclChar, r, cm.comm, ctx->stream));
        if (sc.values_bytes) FG_NCCL(nccl().Send(sc.values(), size_t(sc.values_bytes), ncclChar, r, cm.comm, ctx->stream));
        int64_t nb = byte_base[u][r + 1] - byte_base[u][r];
        if (nb) FG_NCCL(nccl().Recv(static_cast<char*>(oc.data->ptr) + byte_base[u][r], size_t(nb), ncclChar, r, cm.comm, ctx->stream));
      }
    }
  }
  FG_NCCL(nccl().GroupEnd());
  for (size_t u = 0; u < utf8_cols.size(); ++u) {
    Column& oc = out->cols[utf8_cols[u]];
    for (int s = 0; s < W; ++s) {
      if (recv_rows[s] && byte_base[u][s]) {
        int blocks = int(std::min<int64_t>((recv_rows[s] + 255) / 256, int64_t(ctx->sm_count) * 8));
        {
          LaunchTimer lt(ctx, "add_offset_kernel");
          add_offset_kernel<<<blocks, 256, 0, ctx->stream>>>(oc.offsets->as<int32_t>() + row_base[s], recv_rows[s], int32_t(byte_base[u][s]));
        }
        count_launch(ctx);
      }
    }
    {
      LaunchTimer lt(ctx, "store_i32_kernel");
      store_i32_kernel<<<1, 1, 0, ctx->stream>>>(oc.offsets->as<int32_t>() + total_rows, int32_t(byte_base[u][W]));
    }
    count_launch(ctx);
  }
  // the send buffers (parts) must stay alive until the exchange has run
  FG_CUDA(cudaStreamSynchronize(ctx->stream));
  return out;
}

}  // namespace fg

using namespace fg;

extern "C" {

int flockgpu_comm_unique_id(uint8_t out_id[FLOCKGPU_UNIQUE_ID_BYTES]) {
  return guarded([&] {
    FG_CHECK(out_id, FLOCKGPU_ERR_INVALID, "comm_unique_id: null output");
    comm_unique_id(out_id);
  });
}

int flockgpu_comm_init(flockgpu_ctx* ctx, const uint8_t id[FLOCKGPU_UNIQUE_ID_BYTES], int32_t rank, int32_t world_size) {
  return guarded([&] {
    auto c = core_of(ctx);
    FG_CHECK(id, FLOCKGPU_ERR_INVALID, "comm_init: null id");
    std::lock_guard<std::recursive_mutex> g(c->mu);
    FG_CUDA(cudaSetDevice(c->device));
    comm_init(c, id, rank, world_size);
  });
}

int flockgpu_comm_rank(flockgpu_ctx* ctx, int32_t* rank, int32_t* world_size) {
  return guarded([&] {
    auto c = core_of(ctx);
    if (rank) *rank = c->comm ? c->comm->rank : 0;
    if (world_size) *world_size = c->comm ? c->comm->world : 1;
  });
}

int flockgpu_all_to_all(flockgpu_ctx* ctx, flockgpu_table* const* parts, int32_t n_parts, flockgpu_table** out) {
  return guarded([&] {
    auto c = core_of(ctx);
    FG_CHECK(out && parts && n_parts > 0, FLOCKGPU_ERR_INVALID, "all_to_all: bad arguments");
    std::lock_guard<std::recursive_mutex> g(c->mu);
    FG_CUDA(cudaSetDevice(c->device));
    std::vector<TablePtr> ps;
    for (int i = 0; i < n_parts; ++i) {
      FG_CHECK(parts[i] && parts[i]->table, FLOCKGPU_ERR_INVALID, "all_to_all: null partition");
      ps.push_back(parts[i]->table);
    }
    *out = wrap_table(all_to_all(c, ps));
  });
}

}  // extern "C"
