// partition.cu -- K2: RepartitionExec: partitioning = Hash([keys], n).
//
// Reference operator (DataFusion fork; used at flock/src/distributed_plan/planner.rs:153, :160, :223,
// :234; standalone helper flock/src/transmute.rs:77-109; the call shape is spelled out in
// playground/src/distributed_plan/shuffle_writer.rs:105-146): create_hashes(key arrays) ->
// partition = hash % n -> per-partition index lists -> arrow `take` of every column.  Rows keep their
// input order inside a partition.
//
//   partition_ids_kernel     pid[row] = high 32 hash bits scaled to [0, n)  (tables use the LOW bits)
//   partition_select_kernel  one stable single-pass compaction per partition (compact.cuh) appending
//                            the row indices of partition p behind those of partitions < p (the pass's
//                            last tile writes where the next partition starts)
//   gather.cu                materialises each partition's columns from its slice of the index vector
#include <algorithm>

#include "compact.cuh"
#include "internal.h"
#include "rowkeys.cuh"

namespace fg {

struct PartIdArgs {
  int64_t n_rows;
  int32_t packed;
  int32_t n_parts;
  KeyPack pack;
  RowKeys rk;
  ColRef cols[MAX_IN_COLS];
  uint8_t* pid;
};

__host__ __device__ __forceinline__ uint32_t partition_of(uint64_t hash, uint32_t n_parts) {
  return uint32_t(((hash >> 32) * uint64_t(n_parts)) >> 32);
}

__global__ void __launch_bounds__(256) partition_ids_kernel(const __grid_constant__ PartIdArgs a) {
  for (int64_t row = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; row < a.n_rows; row += int64_t(gridDim.x) * blockDim.x) {
    unsigned long long h = a.packed ? fmix64(pack_key(a.pack, a.cols, row)) : hash_row(a.rk, a.cols, row);
    a.pid[row] = uint8_t(partition_of(h, uint32_t(a.n_parts)));
  }
}

struct PartSelectArgs {
  CompactScratch sc;
  const uint8_t* pid;
  int64_t n_rows;
  int32_t part;
  int32_t pad;
  unsigned long long* bases;  // bases[p] = first index slot of partition p; the kernel writes bases[p + 1]
  uint32_t* idx;
};

__global__ void __launch_bounds__(CP_THREADS) partition_select_kernel(const __grid_constant__ PartSelectArgs a) {
  constexpr int E = 4;
  constexpr int CP_ITEMS = 16;
  constexpr int CP_TILE = CP_THREADS * CP_ITEMS;
  __shared__ CompactSmem<E, CP_ITEMS> sm;
  const int tid = threadIdx.x;
  const unsigned long long base = a.bases[a.part];
  long long tile;
  for (int it = 0; (tile = cp_next_tile(sm, a.sc, it)) >= 0; ++it) {
    const int64_t tile_base = tile * CP_TILE;
    unsigned long long bits = 0;
#pragma unroll
    for (int g = 0; g < CP_ITEMS / E; ++g) {
      const int64_t r0 = tile_base + (int64_t(g) * CP_THREADS + tid) * E;
      uint32_t w = 0xffffffffu;
      if (r0 + 3 < a.n_rows) {
        w = *reinterpret_cast<const uint32_t*>(a.pid + r0);
      } else {
        for (int e = 0; e < 4; ++e)
          if (r0 + e < a.n_rows) w = (w & ~(0xffu << (8 * e))) | (uint32_t(a.pid[r0 + e]) << (8 * e));
          else w = (w & ~(0xffu << (8 * e))) | (0xffu << (8 * e));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        bool in_range = r0 + e < a.n_rows;
        bits |= (unsigned long long)(in_range && ((w >> (8 * e)) & 0xffu) == uint32_t(a.part)) << (g * E + e);
      }
    }
    unsigned lane_prefix[CP_ITEMS / E];
    cp_rank_tile<E, CP_ITEMS>(sm, a.sc, tile, bits, lane_prefix);
    // the last tile knows the partition's size: the next pass starts behind it (no separate "advance" launch)
    if (tile == a.sc.num_tiles - 1 && tid == 0) a.bases[a.part + 1] = base + sm.excl + sm.tile_total;
    if (bits && sm.tile_total) {
      unsigned long long m = bits;
      while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        a.idx[base + (unsigned long long)cp_position<E, CP_ITEMS>(sm, bits, k, lane_prefix)] = uint32_t(tile_base + cp_item_index<E>(k, tid));
      }
    }
    __syncthreads();
  }
}

std::vector<TablePtr> hash_partition(const CtxPtr& ctx, const TablePtr& in_ptr, const std::vector<int>& keys, int n_parts) {
  const Table& in = *in_ptr;
  in.dense();
  FG_CHECK(n_parts >= 1 && n_parts <= 255, FLOCKGPU_ERR_INVALID, "hash_partition: n_parts must be in [1, 255], got %d", n_parts);
  FG_CHECK(!keys.empty() && keys.size() <= size_t(MAX_KEY_COLS), FLOCKGPU_ERR_INVALID, "hash_partition: 1..%d key columns", MAX_KEY_COLS);
  FG_CHECK(in.cols.size() <= size_t(MAX_IN_COLS), FLOCKGPU_ERR_UNSUPPORTED, "hash_partition: more than %d columns", MAX_IN_COLS);
  std::vector<int> widths;
  for (int k : keys) {
    FG_CHECK(k >= 0 && k < int(in.cols.size()), FLOCKGPU_ERR_INVALID, "hash_partition: key column %d out of range", k);
    FG_CHECK(!in.cols[k].all_null, FLOCKGPU_ERR_UNSUPPORTED, "hash_partition: NULL key column");
    widths.push_back(in.cols[k].width());
  }
  std::vector<TablePtr> out;
  if (n_parts == 1) {
    out.push_back(in_ptr);
    return out;
  }
  if (in.num_rows == 0) {
    for (int p = 0; p < n_parts; ++p) out.push_back(empty_like(ctx, in));
    return out;
  }
  const int64_t n = in.num_rows;
  BufferPtr pid = alloc(ctx, size_t(n) + 16);
  PartIdArgs ia{};
  ia.n_rows = n;
  ia.packed = keys_packable(widths.data(), int(widths.size())) ? 1 : 0;
  ia.n_parts = n_parts;
  for (size_t i = 0; i < in.cols.size(); ++i) {
    ia.cols[i].data = in.cols[i].values();
    ia.cols[i].offsets = in.cols[i].offs();
    ia.cols[i].dtype = in.cols[i].dtype;
  }
  ia.rk.n = int(keys.size());
  for (size_t i = 0; i < keys.size(); ++i) ia.rk.col[i] = keys[i];
  if (ia.packed) {
    ia.pack.n = int(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) {
      ia.pack.col[i] = keys[i];
      ia.pack.width[i] = widths[i];
    }
  }
  ia.pid = pid->as<uint8_t>();
  int grid = int(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, int64_t(ctx->sm_count) * 8)));
  {
    LaunchTimer lt(ctx, "partition_ids_kernel");
    partition_ids_kernel<<<grid, 256, 0, ctx->stream>>>(ia);
  }
  FG_CUDA(cudaGetLastError());
  count_launch(ctx);

  // bases live in d_scalars[16 .. 16 + n_parts]; the per-pass survivor count in d_scalars[5]
  FG_CHECK(16 + n_parts + 1 <= 384, FLOCKGPU_ERR_UNSUPPORTED, "hash_partition: too many partitions");  // d_scalars[384 ..] belongs to gather.cu
  unsigned long long* bases = ctx->d_scalars + 16;
  FG_CUDA(cudaMemsetAsync(bases, 0, sizeof(unsigned long long) * (n_parts + 1), ctx->stream));
  BufferPtr idx = alloc(ctx, size_t(n) * 4);
  PartSelectArgs sa{};
  const long long tiles = (n + CP_THREADS * 16 - 1) / (CP_THREADS * 16);
  sa.pid = pid->as<uint8_t>();
  sa.n_rows = n;
  sa.bases = bases;
  sa.idx = idx->as<uint32_t>();
  int per_sm = 1;
  FG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, partition_select_kernel, CP_THREADS, 0));
  const long long resident = (long long)ctx->sm_count * std::max(per_sm, 1);
  for (int p = 0; p < n_parts; ++p) {
    sa.part = p;
    sa.sc = prepare_compact(ctx, tiles, resident, ctx->d_scalars + 5);
    {
      LaunchTimer lt(ctx, "partition_select_kernel");
      launch_compact(ctx, partition_select_kernel, sa.sc, sa);
    }
    FG_CUDA(cudaGetLastError());
    count_launch(ctx);
  }
  std::vector<unsigned long long> hb(n_parts + 1);
  read_scalars(ctx, 16, n_parts + 1, hb.data());
  FG_CHECK(int64_t(hb[n_parts]) == n, FLOCKGPU_ERR_CUDA, "hash_partition: partition sizes sum to %llu, expected %lld", hb[n_parts], (long long)n);
  std::vector<int> all_cols;
  for (size_t i = 0; i < in.cols.size(); ++i) all_cols.push_back(int(i));
  for (int p = 0; p < n_parts; ++p) {
    int64_t cnt = int64_t(hb[p + 1] - hb[p]);
    out.push_back(gather_rows(ctx, in, all_cols, idx->as<uint32_t>() + hb[p], cnt));
  }
  // `idx` must outlive the gathers: they are stream-ordered before its (stream-ordered) free
  return out;
}

}  // namespace fg

using namespace fg;

extern "C" int flockgpu_hash_partition(flockgpu_ctx* ctx, const flockgpu_table* in, const int32_t* key_cols, int32_t n_keys, int32_t n_parts,
                                       flockgpu_table** out_parts) {
  return guarded([&] {
    auto c = core_of(ctx);
    FG_CHECK(out_parts && in && in->table && key_cols && n_keys > 0, FLOCKGPU_ERR_INVALID, "hash_partition: null or empty argument");
    std::lock_guard<std::recursive_mutex> g(c->mu);
    FG_CUDA(cudaSetDevice(c->device));
    std::vector<int> keys(key_cols, key_cols + n_keys);
    std::vector<TablePtr> parts = hash_partition(c, in->table, keys, n_parts);
    for (int p = 0; p < n_parts; ++p) out_parts[p] = wrap_table(parts[p]);
  });
}
