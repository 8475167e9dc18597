// filter_project.cu -- K1: FilterExec + CoalesceBatchesExec + ProjectionExec as one pass over HBM.
//
// Reference operators replaced (DataFusion fork, not in tree; shapes evidenced at
// flock/src/distributed_plan/planner.rs:90-92 (q1), :120-124 (q2), :151-163 (q3 filters)):
//   FilterExec:     predicate.evaluate(batch) -> BooleanArray -> arrow filter_record_batch (stable)
//   ProjectionExec: per-batch expression evaluation; plain Column exprs are zero-copy
//
// filter_compact_kernel: ONE launch for the whole relation (all 64 Ki-row batches: a 1.3 MB batch is
// 0.2 us of HBM traffic, far below launch latency, so per-batch launches cannot work).  CTAs take tiles of
// 256 x {4, 16, 32, 64} rows (sized so that the relation is one wave of resident CTAs whenever possible),
// evaluate the predicate (PredI32: one multiply-add + compare per row for the NEXMark shapes; the generic term
// interpreter otherwise), rank the survivors (compact.cuh), obtain the tile's global output offset from the
// grid-wide prefix protocol (each input byte is read once, no second pass), and write the surviving rows of every
// fixed-width output column in input order.  Utf8 outputs leave through the selection vector and gather.cu.
// Host-resident (page-locked, zero-copy fed) predicate columns are read in place over PCIe.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "compact.cuh"
#include "expr_compile.h"
#include "internal.h"
#include "pred_i32.h"

namespace fg {

constexpr int FP_THREADS = CP_THREADS;

struct FilterArgs {
  int64_t n_rows;
  int32_t n_out;
  int32_t pad;
  CompactScratch sc;                 // tickets / prefix scratch of this launch (compact.cuh)
  uint32_t* sel_out;                 // optional: indices of surviving rows
  unsigned long long* out_count;     // total surviving rows
  unsigned long long* host_count;    // pinned host slot that also receives the total (NULL: none)
  int* err_flag;                     // set to 1 on divide-by-zero
  unsigned long long* trace;         // debug (FLOCKGPU_TRACE): 8 globaltimer stamps per tile, else NULL
  ColRef cols[MAX_IN_COLS];
  OutCol outs[MAX_OUT_COLS];
};

// ---- predicate functors ---------------------------------------------------------------------------
// A functor evaluates this thread's FP_ITEMS rows of a tile and returns one bit per item.  E = rows a
// thread loads contiguously (vector width); item k is row
//   tile_base + ((k / E) * FP_THREADS + tid) * E + k % E.
// ROWS = rows per thread per tile: 16 for large inputs; 4 for small ones, where 4 Ki-row tiles would leave most SMs
// idle and every thread would walk 16 rows' worth of dependent loads (NEXMark q3: 200 K persons = 49 tiles)
template <int ROWS>
struct PredGenericT {
  static constexpr int E = 1;
  static constexpr int I = ROWS;
  static constexpr int MIN_CTAS = 1;
  Predicate p;
  __device__ __forceinline__ unsigned long long eval(const ColRef* cols, int64_t tile_base, int64_t n_rows, int tid, int* err) const {
    int64_t rows[I];
#pragma unroll
    for (int k = 0; k < I; ++k) {
      int64_t r = tile_base + int64_t(k) * FP_THREADS + tid;
      rows[k] = r < n_rows ? r : -1;
    }
    return eval_predicate<I>(p, cols, rows, err);
  }
};


template <int MODE, int ITEMS>
struct PredI32 {
  static constexpr int E = 4;
  static constexpr int I = ITEMS;  // rows per thread per tile
  static constexpr int MIN_CTAS = 5;  // <= 51 registers: 740 resident CTAs, so 10 M rows (611 tiles) are one wave
  const int32_t* col;
  PredI32Consts k;
  const void* const* chunks;  // host-resident column (zero-copy feed): chunk bases, else NULL
  int32_t chunk_shift;
  // address of rows [row0, row0 + 4): row0 is a multiple of 4 and chunk lengths are powers of two >= 4096
  __device__ __forceinline__ const int32_t* at(int64_t row0) const {
    if (!chunks) return col + row0;
    return static_cast<const int32_t*>(chunks[row0 >> chunk_shift]) + (row0 & ((int64_t(1) << chunk_shift) - 1));
  }
  __device__ __forceinline__ bool test(int32_t x) const { return pred_i32_test<MODE>(k, x); }
  __device__ __forceinline__ unsigned test4(const int4& v) const {
    return unsigned(test(v.x)) | (unsigned(test(v.y)) << 1) | (unsigned(test(v.z)) << 2) | (unsigned(test(v.w)) << 3);
  }
  // Streams the thread's rows in rounds of four 16-byte loads: only the one-bit verdicts are kept (the values of
  // the ~1 % survivors are re-read from L2 when they are written), so the kernel needs ~48 registers and five CTAs
  // stay resident per SM -- a 10 M-row relation (611 tiles of 16 Ki rows) is then a single wave.
  __device__ __forceinline__ unsigned long long eval(const ColRef*, int64_t tile_base, int64_t n_rows, int tid, int*) const {
    constexpr int L = 4;  // independent 16-byte loads in flight per thread and round (8 costs 80 registers = 3 CTAs/SM: no single wave)
    constexpr int ROUNDS = I / (4 * L);
    if (!chunks && tile_base + int64_t(FP_THREADS) * I <= n_rows) {
      // full tile of a device-resident column (every tile but the last): no bounds, no chunk lookup
      const int4* p = reinterpret_cast<const int4*>(col + tile_base) + tid;
      unsigned half[2] = {0u, 0u};  // verdicts of rounds {0, 1} and {2, 3}: 32-bit shifts only
#pragma unroll
      for (int h = 0; h < (ROUNDS + 1) / 2; ++h) {
#pragma unroll 1
        for (int c = 2 * h; c < ROUNDS && c < 2 * h + 2; ++c) {
          int4 v[L];
#pragma unroll
          for (int j = 0; j < L; ++j) v[j] = ldg_stream_v4(p + (c * L + j) * FP_THREADS);
          unsigned b16 = 0;
#pragma unroll
          for (int j = 0; j < L; ++j) b16 |= test4(v[j]) << (4 * j);
          half[h] |= b16 << (16 * (c & 1));
        }
      }
      return (unsigned long long)half[0] | ((unsigned long long)half[1] << 32);
    }
    unsigned long long bits = 0;
#pragma unroll 1
    for (int c = 0; c < ROUNDS; ++c) {
      int4 v[L];
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int64_t row0 = tile_base + (int64_t(c * L + j) * FP_THREADS + tid) * 4;
        if (row0 + 3 < n_rows) {
          v[j] = ldg_stream_v4(at(row0));
        } else {
          v[j] = make_int4(0, 0, 0, 0);
          if (row0 + 0 < n_rows) v[j].x = at(row0)[0];
          if (row0 + 1 < n_rows) v[j].y = at(row0)[1];
          if (row0 + 2 < n_rows) v[j].z = at(row0)[2];
        }
      }
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int64_t row0 = tile_base + (int64_t(c * L + j) * FP_THREADS + tid) * 4;
        unsigned b = test4(v[j]);
        const int64_t left = n_rows - row0;  // mask rows past the end
        if (left < 4) b &= left <= 0 ? 0u : ((1u << left) - 1u);
        bits |= (unsigned long long)b << (4 * (c * L + j));
      }
    }
    return bits;
  }
};

__device__ __forceinline__ void copy_value(void* dst, const void* src, int width, int64_t pos, int64_t row) {
  if (width == 4) static_cast<uint32_t*>(dst)[pos] = static_cast<const uint32_t*>(src)[row];
  else static_cast<uint64_t*>(dst)[pos] = static_cast<const uint64_t*>(src)[row];
}

template <class PredFn>
__global__ void __launch_bounds__(FP_THREADS, PredFn::MIN_CTAS) filter_compact_kernel(const __grid_constant__ PredFn pred, const __grid_constant__ FilterArgs a) {
  constexpr int E = PredFn::E;
  constexpr int I = PredFn::I;
  constexpr int TILE = FP_THREADS * I;
  __shared__ CompactSmem<E, I> sm;
  const int tid = threadIdx.x;
  const CompactScratch& sc = a.sc;
  int err = 0;

  auto stamp = [&](long long tile, int k) {
    if (a.trace && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[tile * 8 + k] = t;
    }
  };
  unsigned long long t_begin = 0;
  if (a.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
  long long tile;
  for (int it = 0; (tile = cp_next_tile(sm, sc, it)) >= 0; ++it) {
    const int64_t tile_base = tile * TILE;
    if (a.trace && tid == 0) a.trace[tile * 8 + 0] = t_begin;
    stamp(tile, 1);
    const unsigned long long bits = pred.eval(a.cols, tile_base, a.n_rows, tid, &err);
    stamp(tile, 2);
    unsigned lane_prefix[I / E];
    // While the CTA waits for the grid-wide prefix, pull the survivors' pass-through values (one dependent ~1 us
    // DRAM read each) into L2: the write phase below then finds them there.
    auto prefetch_survivors = [&] {
      if (!bits) return;
      for (int c = 0; c < a.n_out; ++c) {
        const OutCol& oc = a.outs[c];
        if (oc.kind != OUT_PASS) continue;
        const ColRef& src = a.cols[oc.src_col];
        if (src.chunks) continue;  // host-resident: fetched once, when written
        unsigned long long m = bits;
#pragma unroll 1
        for (int q = 0; q < 4 && m; ++q) {
          const int k = __ffsll((long long)m) - 1;
          m &= m - 1;
          prefetch_l2(static_cast<const char*>(src.data) + (tile_base + cp_item_index<E>(k, tid)) * oc.width);
        }
      }
    };
    cp_rank_tile<E, I>(sm, sc, tile, bits, lane_prefix, prefetch_survivors);
    stamp(tile, 3);

    // ---- write survivors in input order
    // A thread may own many survivors (NEXMark's hot auction id: when it satisfies the predicate half of the rows
    // of a stretch survive), and every pass-through value is a dependent ~1 us DRAM read: gather four survivors
    // at a time so that their loads are in flight together.
    if (bits && sm.tile_total) {
      unsigned long long m = bits;
      while (m) {
        int64_t pos[4], row[4];
        int nb = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          pos[q] = 0;
          row[q] = -1;
          if (m) {
            const int k = __ffsll((long long)m) - 1;
            m &= m - 1;
            pos[q] = cp_position<E, I>(sm, bits, k, lane_prefix);
            row[q] = tile_base + cp_item_index<E>(k, tid);
            nb = q + 1;
          }
        }
        if (a.sel_out) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < nb) a.sel_out[pos[q]] = uint32_t(row[q]);
        }
        for (int c = 0; c < a.n_out; ++c) {
          const OutCol& oc = a.outs[c];
          if (oc.kind == OUT_PASS) {
            const ColRef& sc_col = a.cols[oc.src_col];
            const void* src = sc_col.data;
            if (sc_col.chunks) {
              // host-resident source: one PCIe read per survivor
              const int64_t mask = (int64_t(1) << sc_col.chunk_shift) - 1;
              uint64_t v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[q] = 0;
                if (q < nb) {
                  const char* base = static_cast<const char*>(sc_col.chunks[row[q] >> sc_col.chunk_shift]) + (row[q] & mask) * oc.width;
                  v[q] = oc.width == 4 ? uint64_t(*reinterpret_cast<const uint32_t*>(base)) : *reinterpret_cast<const uint64_t*>(base);
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < nb) {
                  if (oc.width == 4) static_cast<uint32_t*>(oc.dst)[pos[q]] = uint32_t(v[q]);
                  else static_cast<uint64_t*>(oc.dst)[pos[q]] = v[q];
                }
            } else if (oc.width == 4) {
              uint32_t v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = q < nb ? static_cast<const uint32_t*>(src)[row[q]] : 0u;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < nb) static_cast<uint32_t*>(oc.dst)[pos[q]] = v[q];
            } else {
              uint64_t v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = q < nb ? static_cast<const uint64_t*>(src)[row[q]] : 0ull;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < nb) static_cast<uint64_t*>(oc.dst)[pos[q]] = v[q];
            }
          } else {
            Val acc[4];
            eval_chain<4>(oc.chain, a.cols, row, acc, &err);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nb) store_val(oc.dst, oc.out_dtype, pos[q], acc[q]);
          }
        }
      }
    }
    __syncthreads();  // sm is reused by the next tile
    stamp(tile, 4);
    if (a.trace && tid == 0) a.trace[tile * 8 + 6] = blockIdx.x;
    if (a.trace && tid == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.trace[tile * 8 + 7] = smid;
    }
  }
  if (err) *a.err_flag = 1;
  if (a.trace && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(long long)blockIdx.x * 8 + 5] = t;   // per CTA (indexed by block, not tile)
  }
}

// ---- pure projection (no predicate): streaming evaluation of the computed columns ----------------
struct ProjectArgs {
  int64_t n_rows;
  int32_t n_out;
  int32_t pad;
  int* err_flag;
  ColRef cols[MAX_IN_COLS];
  OutCol outs[MAX_OUT_COLS];
};

constexpr int PJ_THREADS = 256;
constexpr int PJ_ROWS = 4;

__global__ void __launch_bounds__(PJ_THREADS) project_generic_kernel(const __grid_constant__ ProjectArgs a) {
  int err = 0;
  const int64_t stride = int64_t(gridDim.x) * PJ_THREADS * PJ_ROWS;
  for (int64_t base = int64_t(blockIdx.x) * PJ_THREADS * PJ_ROWS; base < a.n_rows; base += stride) {
    int64_t rows[PJ_ROWS];
#pragma unroll
    for (int j = 0; j < PJ_ROWS; ++j) {
      int64_t r = base + int64_t(j) * PJ_THREADS + threadIdx.x;
      rows[j] = r < a.n_rows ? r : -1;
    }
    for (int c = 0; c < a.n_out; ++c) {
      const OutCol& oc = a.outs[c];
      Val acc[PJ_ROWS];
      eval_chain<PJ_ROWS>(oc.chain, a.cols, rows, acc, &err);
#pragma unroll
      for (int j = 0; j < PJ_ROWS; ++j)
        if (rows[j] >= 0) store_val(oc.dst, oc.out_dtype, rows[j], acc[j]);
    }
  }
  if (err) *a.err_flag = 1;
}

// NEXMark q1: price' = lit * CAST(price AS Float64).  One IEEE multiply per row (__dmul_rn: the i32 ->
// f64 conversion is exact, so the product has a single rounding).
// 16 B in, 32 B out per thread-iteration; 4 B + 8 B per row is all the traffic there is.
__global__ void __launch_bounds__(PJ_THREADS) project_i32_to_f64_mul_kernel(const int32_t* __restrict__ in, double* __restrict__ out,
                                                                            int64_t n_rows, double lit) {
  const int64_t n4 = n_rows >> 2;
  const int64_t stride = int64_t(gridDim.x) * PJ_THREADS;
  for (int64_t i = int64_t(blockIdx.x) * PJ_THREADS + threadIdx.x; i < n4; i += stride) {
    int4 v = ldg_stream_v4(in + i * 4);
    stg_stream_v2f64(out + i * 4, __dmul_rn(lit, double(v.x)), __dmul_rn(lit, double(v.y)));
    stg_stream_v2f64(out + i * 4 + 2, __dmul_rn(lit, double(v.z)), __dmul_rn(lit, double(v.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n_rows & 3)) {
    int64_t r = (n4 << 2) + threadIdx.x;
    out[r] = __dmul_rn(lit, double(in[r]));
  }
}

__global__ void __launch_bounds__(256) and_validity_kernel(const uint8_t* a, const uint8_t* b, const uint8_t* c, const uint8_t* d, uint8_t* out, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    uint8_t v = a[i];
    if (b) v &= b[i];
    if (c) v &= c[i];
    if (d) v &= d[i];
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static std::vector<ColInfo> col_infos(const Table& t) {
  std::vector<ColInfo> v;
  for (const Column& c : t.cols) v.push_back({c.dtype, c.name, c.format, c.validity != nullptr});
  return v;
}

static void fill_colrefs(const Table& t, ColRef* refs) {
  FG_CHECK(t.cols.size() <= size_t(MAX_IN_COLS), FLOCKGPU_ERR_UNSUPPORTED, "filter_project: more than %d input columns", MAX_IN_COLS);
  for (size_t i = 0; i < t.cols.size(); ++i) {
    FG_CHECK(!t.cols[i].all_null, FLOCKGPU_ERR_UNSUPPORTED, "filter_project: NULL input column \"%s\"", t.cols[i].name.c_str());
    refs[i].data = t.cols[i].values();
    refs[i].offsets = t.cols[i].offs();
    refs[i].dtype = t.cols[i].dtype;
    refs[i].chunk_shift = t.cols[i].chunks ? t.cols[i].chunks->shift : 0;
    refs[i].chunks = t.cols[i].chunks ? static_cast<const void* const*>(t.cols[i].chunks->table->ptr) : nullptr;
    refs[i].validity = t.cols[i].valid();
  }
}

template <class PredFn>
static void launch_filter(const CtxPtr& ctx, const PredFn& pred, FilterArgs args) {
  constexpr int64_t TILE = int64_t(FP_THREADS) * PredFn::I;
  const int64_t num_tiles = (args.n_rows + TILE - 1) / TILE;
  const void* k = reinterpret_cast<const void*>(&filter_compact_kernel<PredFn>);
  args.sc = prepare_compact(ctx, num_tiles, resident_ctas(ctx, k, FP_THREADS), args.out_count);
  const int grid = args.sc.grid;
  args.sc.host_count = args.host_count;
  BufferPtr trace_buf;
  static const char* trace_path = getenv("FLOCKGPU_TRACE");
  if (trace_path) {
    trace_buf = alloc(ctx, size_t(num_tiles + 2048) * 64);
    FG_CUDA(cudaMemsetAsync(trace_buf->ptr, 0, size_t(num_tiles + 2048) * 64, ctx->stream));
    args.trace = trace_buf->as<unsigned long long>();
  }
  {
    LaunchTimer lt(ctx, "filter_compact_kernel");
    launch_compact(ctx, filter_compact_kernel<PredFn>, args.sc, pred, args);
  }
  FG_CUDA(cudaGetLastError());
  count_launch(ctx);
  if (trace_path) {
    // debug only: dump the per-tile time stamps of this launch (overwrites the file each launch)
    std::vector<unsigned long long> h(size_t(num_tiles + 2048) * 8);
    FG_CUDA(cudaMemcpyAsync(h.data(), trace_buf->ptr, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (FILE* f = fopen(trace_path, "w")) {
      fprintf(f, "%lld %d\n", (long long)num_tiles, grid);
      for (size_t i = 0; i < h.size() / 8; ++i) {
        for (int j = 0; j < 8; ++j) fprintf(f, "%llu ", h[i * 8 + j]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
}

TablePtr filter_project(const CtxPtr& ctx, const TablePtr& in_ptr, const Expr* predicate, const std::vector<Expr>& projections,
                        const std::vector<std::string>& names) {
  const Table& in = *in_ptr;
  // a pure column selection / renaming of a relation that is still in table form (DeferredTable) keeps it there
  if (in_ptr->deferred && !predicate && !projections.empty()) {
    std::vector<ColInfo> dinfos = col_infos(in);
    std::vector<int> src;
    std::vector<std::string> out_names;
    bool plain = true;
    for (size_t i = 0; i < projections.size() && plain; ++i) {
      CompiledValue v = compile_value(projections[i], dinfos);
      plain = v.passthrough;
      src.push_back(v.src_col);
      out_names.push_back(i < names.size() && !names[i].empty() ? names[i] : (plain ? in.cols[v.src_col].name : std::string()));
    }
    if (plain) {
      std::shared_ptr<DeferredTable> d = in_ptr->deferred;
      if (TablePtr t = d->project(in, src, out_names)) return t;
    }
  }
  in.resolve();
  std::vector<ColInfo> infos = col_infos(in);
  // ---- compile
  std::vector<CompiledValue> vals;
  if (projections.empty()) {
    for (size_t i = 0; i < in.cols.size(); ++i) {
      CompiledValue v;
      v.passthrough = true;
      v.src_col = int(i);
      v.dtype = in.cols[i].dtype;
      v.format = in.cols[i].format;
      vals.push_back(v);
    }
  } else {
    for (const Expr& e : projections) vals.push_back(compile_value(e, infos));
  }
  auto out_name = [&](size_t i) -> std::string {
    if (i < names.size() && !names[i].empty()) return names[i];
    if (vals[i].passthrough) return in.cols[vals[i].src_col].name;
    return projections.empty() ? std::string() : expr_to_string(projections[i], infos);
  };

  auto out = std::make_shared<Table>();
  out->ctx = ctx;
  out->metadata = in.metadata;
  // rows stay on their rank: the routing property of an exchanged input survives when its columns pass through by name
  if (!in.partitioned_on.empty()) {
    bool kept = true;
    for (const std::string& p : in.partitioned_on) {
      bool found = false;
      for (size_t i = 0; i < vals.size(); ++i) found |= vals[i].passthrough && in.cols[vals[i].src_col].name == p && out_name(i) == p;
      kept &= found;
    }
    if (kept) {
      out->partitioned_on = in.partitioned_on;
      out->partition_world = in.partition_world;
    }
  }

  int* err_flag = reinterpret_cast<int*>(ctx->d_scalars + 8);
  bool check_err = false;

  if (!predicate) {
    // ------------------------------------------------------------------ ProjectionExec only
    in.dense();
    out->num_rows = in.num_rows;
    ProjectArgs pa{};
    pa.n_rows = in.num_rows;
    pa.err_flag = err_flag;
    fill_colrefs(in, pa.cols);
    for (size_t i = 0; i < vals.size(); ++i) {
      const CompiledValue& v = vals[i];
      Column c;
      c.name = out_name(i);
      c.dtype = v.dtype;
      c.format = v.format;
      c.length = in.num_rows;
      if (v.passthrough) {
        const Column& s = in.cols[v.src_col];
        c.nullable = s.nullable;
        c.data = s.data;  // zero-copy, like the Arc clone in ProjectionExec
        c.offsets = s.offsets;
        c.values_bytes = s.values_bytes;
        c.all_null = s.all_null;
        c.validity = s.validity;
      } else {
        // a computed value is NULL where any column it reads is NULL
        std::vector<const uint8_t*> srcs;
        auto note = [&](int col) {
          if (col >= 0 && in.cols[col].validity) srcs.push_back(in.cols[col].valid());
        };
        note(v.chain.start_col);
        for (int st = 0; st < v.chain.n_steps; ++st)
          if (v.chain.steps[st].src == 1) note(v.chain.steps[st].col);
        if (!srcs.empty() && in.num_rows > 0) {
          FG_CHECK(srcs.size() <= 4, FLOCKGPU_ERR_UNSUPPORTED, "projection: an expression over more than 4 columns with NULLs");
          c.validity = alloc(ctx, size_t(in.num_rows));
          c.nullable = true;
          int grid = int(std::min<int64_t>((in.num_rows + 255) / 256, int64_t(ctx->sm_count) * 8));
          {
            LaunchTimer lt(ctx, "and_validity_kernel");
            and_validity_kernel<<<grid, 256, 0, ctx->stream>>>(srcs[0], srcs.size() > 1 ? srcs[1] : nullptr, srcs.size() > 2 ? srcs[2] : nullptr,
                                                           srcs.size() > 3 ? srcs[3] : nullptr, c.validity->as<uint8_t>(), in.num_rows);
          }
          FG_CUDA(cudaGetLastError());
          count_launch(ctx);
        }
        c.data = alloc(ctx, size_t(in.num_rows) * dtype_width(v.dtype));
        if (in.num_rows > 0) {
          if (v.fast == FAST_VAL_I32_TO_F64_MUL) {
            int64_t n4 = std::max<int64_t>(1, in.num_rows / 4);
            int grid = int(std::min<int64_t>((n4 + PJ_THREADS - 1) / PJ_THREADS, int64_t(ctx->sm_count) * 8));
            {
              LaunchTimer lt(ctx, "project_i32_to_f64_mul_kernel");
              project_i32_to_f64_mul_kernel<<<grid, PJ_THREADS, 0, ctx->stream>>>(
                static_cast<const int32_t*>(in.cols[v.chain.start_col].values()), c.data->as<double>(), in.num_rows, v.fast_lit);
            }
            FG_CUDA(cudaGetLastError());
            count_launch(ctx);
          } else {
            FG_CHECK(pa.n_out < MAX_OUT_COLS, FLOCKGPU_ERR_UNSUPPORTED, "projection: more than %d computed columns", MAX_OUT_COLS);
            OutCol& oc = pa.outs[pa.n_out++];
            oc.kind = OUT_COMPUTED;
            oc.src_col = -1;
            oc.out_dtype = v.dtype;
            oc.width = dtype_width(v.dtype);
            oc.dst = c.data->ptr;
            oc.chain = v.chain;
            check_err |= v.has_div_by_col;
          }
        }
      }
      out->cols.push_back(std::move(c));
    }
    if (pa.n_out > 0 && in.num_rows > 0) {
      int64_t per_block = int64_t(PJ_THREADS) * PJ_ROWS;
      int grid = int(std::min<int64_t>((in.num_rows + per_block - 1) / per_block, int64_t(ctx->sm_count) * 8));
      if (check_err) FG_CUDA(cudaMemsetAsync(err_flag, 0, sizeof(int), ctx->stream));
      {
        LaunchTimer lt(ctx, "project_generic_kernel");
        project_generic_kernel<<<grid, PJ_THREADS, 0, ctx->stream>>>(pa);
      }
      FG_CUDA(cudaGetLastError());
      count_launch(ctx);
      if (check_err) {
        unsigned long long flag = 0;
        read_scalars(ctx, 8, 1, &flag);
        FG_CHECK((flag & 0xffffffffull) == 0, FLOCKGPU_ERR_EXECUTION, "Divide by zero");
      }
    }
    return out;
  }

  // -------------------------------------------------------------------- FilterExec (+ projection)
  CompiledPredicate cp = compile_predicate(*predicate, infos);
  check_err = cp.prog.has_div_by_col != 0;
  // host-resident (zero-copy) columns can be consumed in place only by the vectorised predicate with plain
  // pass-through fixed-width outputs; anything else scans the relation more than once and copies it to HBM first
  if (in.has_host_columns()) {
    bool in_place = cp.fast.kind != FAST_PRED_NONE;
    for (const CompiledValue& v : vals) in_place &= v.passthrough && v.dtype != FLOCKGPU_UTF8;
    if (!in_place) in.dense();
  }
  if (in.num_rows == 0) {
    // nothing to scan: an empty table with the output schema
    out->num_rows = 0;
    for (size_t i = 0; i < vals.size(); ++i) {
      Column c;
      c.name = out_name(i);
      c.dtype = vals[i].dtype;
      c.format = vals[i].format;
      c.length = 0;
      c.data = alloc(ctx, 0);
      if (c.dtype == FLOCKGPU_UTF8) {
        c.offsets = alloc(ctx, 4);
        FG_CUDA(cudaMemsetAsync(c.offsets->ptr, 0, 4, ctx->stream));
      }
      out->cols.push_back(std::move(c));
    }
    return out;
  }

  FilterArgs fa{};
  fa.n_rows = in.num_rows;
  fa.out_count = ctx->d_scalars + 0;
  fa.err_flag = err_flag;
  fill_colrefs(in, fa.cols);

  bool need_sel = false;
  std::vector<int> utf8_outs;
  for (size_t i = 0; i < vals.size(); ++i) {
    const CompiledValue& v = vals[i];
    Column c;
    c.name = out_name(i);
    c.dtype = v.dtype;
    c.format = v.format;
    if (v.passthrough) c.nullable = in.cols[v.src_col].nullable;
    if (!v.passthrough) {
      // a computed output inside the fused filter kernel cannot carry NULLs (no validity output there)
      auto refuse = [&](int col) {
        if (col >= 0) require_no_nulls(in.cols[col], "filter: computed projection");
      };
      refuse(v.chain.start_col);
      for (int st = 0; st < v.chain.n_steps; ++st)
        if (v.chain.steps[st].src == 1) refuse(v.chain.steps[st].col);
    }
    if (v.dtype == FLOCKGPU_UTF8 || (v.passthrough && in.cols[v.src_col].validity)) {
      // Utf8 outputs and outputs with validity bytes are taken through the selection vector behind the kernel
      need_sel = true;
      utf8_outs.push_back(int(i));
    } else {
      FG_CHECK(fa.n_out < MAX_OUT_COLS, FLOCKGPU_ERR_UNSUPPORTED, "filter: more than %d fixed-width output columns", MAX_OUT_COLS);
      c.data = alloc(ctx, size_t(in.num_rows) * dtype_width(v.dtype));  // worst case: every row survives
      OutCol& oc = fa.outs[fa.n_out++];
      oc.kind = v.passthrough ? OUT_PASS : OUT_COMPUTED;
      oc.src_col = v.src_col;
      oc.out_dtype = v.dtype;
      oc.width = dtype_width(v.dtype);
      oc.dst = c.data->ptr;
      oc.chain = v.chain;
      check_err |= v.has_div_by_col;
    }
    out->cols.push_back(std::move(c));
  }
  BufferPtr sel;
  if (need_sel) {
    sel = alloc(ctx, size_t(in.num_rows) * 4);
    fa.sel_out = sel->as<uint32_t>();
  }
  if (check_err) FG_CUDA(cudaMemsetAsync(err_flag, 0, sizeof(int), ctx->stream));
  // fixed-width outputs only: the survivor count stays in flight.  The kernel stores it straight into a pinned host
  // slot (no 8-byte D2H copy on the stream); the first consumer that needs the number waits for the event
  // recorded behind the launch (Table::resolve).
  std::shared_ptr<PendingRows> pending;
  if (!check_err && utf8_outs.empty()) {
    pending = reserve_row_count(ctx);
    fa.host_count = pending->host_slot();
  }
  // zero-copy feed: what this launch reads over the host link (flockgpu_bytes_moved) -- the predicate column whole,
  // the pass-through columns once per survivor
  int64_t zc_bytes_per_survivor = 0;
  if (in.has_host_columns()) {
    if (cp.fast.kind != FAST_PRED_NONE && in.cols[cp.fast.col].chunks)
      ctx->h2d_bytes.fetch_add(in.num_rows * int64_t(in.cols[cp.fast.col].width()), std::memory_order_relaxed);
    for (const CompiledValue& v : vals)
      if (v.passthrough && in.cols[v.src_col].chunks) zc_bytes_per_survivor += in.cols[v.src_col].width();
    if (pending) pending->h2d_bytes_per_row = zc_bytes_per_survivor;
  }

  // rows per thread per tile of the vectorised functors (FLOCKGPU_FILTER_ITEMS = 16 | 32 | 64 overrides, for tuning)
  // The smallest tile that still keeps the relation a single wave (<= ~700 tiles of 256 x items rows): small inputs
  // spread over more SMs, 10 M-row relations use 64.
  static const int items_env = [] {
    const char* e = getenv("FLOCKGPU_FILTER_ITEMS");
    int v = e ? atoi(e) : 0;
    return (v == 16 || v == 32 || v == 64) ? v : 0;
  }();
  const int64_t single_wave_tiles = std::max<int64_t>(64, int64_t(ctx->sm_count) * 5 - 40);
  const int items = items_env ? items_env
                    : in.num_rows <= single_wave_tiles * FP_THREADS * 16 ? 16
                    : in.num_rows <= single_wave_tiles * FP_THREADS * 32 ? 32 : 64;
  auto launch_i32 = [&](int64_t modulus) {
    const Column& pc = in.cols[cp.fast.col];
    const int32_t* col = static_cast<const int32_t*>(pc.values());
    const void* const* chunks = pc.chunks ? static_cast<const void* const*>(pc.chunks->table->ptr) : nullptr;
    const int32_t chunk_shift = pc.chunks ? pc.chunks->shift : 0;
    PredI32Consts k;
    const int mode = pred_i32_consts(modulus, cp.fast.cmp, cp.fast.rhs, &k);
    auto go = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      if (items == 16) launch_filter(ctx, PredI32<MODE, 16>{col, k, chunks, chunk_shift}, fa);
      else if (items == 32) launch_filter(ctx, PredI32<MODE, 32>{col, k, chunks, chunk_shift}, fa);
      else launch_filter(ctx, PredI32<MODE, 64>{col, k, chunks, chunk_shift}, fa);
    };
    if (mode == 0) go(std::integral_constant<int, 0>{});
    else if (mode == 1) go(std::integral_constant<int, 1>{});
    else go(std::integral_constant<int, 2>{});
  };
  switch (cp.fast.kind) {
    case FAST_PRED_I32_MOD_CMP:
      launch_i32(cp.fast.modulus);
      break;
    case FAST_PRED_I32_CMP:
      launch_i32(0);
      break;
    default: {
      if (in.num_rows >= int64_t(ctx->sm_count) * 4 * FP_THREADS * 16) launch_filter(ctx, PredGenericT<16>{cp.prog}, fa);
      else launch_filter(ctx, PredGenericT<4>{cp.prog}, fa);
    }
  }

  if (pending) {
    commit_row_count(pending);
    out->pending = pending;
    out->num_rows = -1;
    for (Column& c : out->cols) c.length = -1;
    return out;
  }
  unsigned long long scalars[9];
  read_scalars(ctx, 0, 9, scalars);
  if (check_err) FG_CHECK((scalars[8] & 0xffffffffull) == 0, FLOCKGPU_ERR_EXECUTION, "Divide by zero");
  int64_t n_sel = int64_t(scalars[0]);
  FG_CHECK(n_sel >= 0 && n_sel <= in.num_rows, FLOCKGPU_ERR_CUDA, "filter: corrupt survivor count %lld", (long long)n_sel);
  ctx->h2d_bytes.fetch_add(n_sel * zc_bytes_per_survivor, std::memory_order_relaxed);
  out->num_rows = n_sel;
  for (Column& c : out->cols) c.length = n_sel;
  if (!utf8_outs.empty()) {
    std::vector<const Column*> src;
    for (int i : utf8_outs) src.push_back(&in.cols[vals[i].src_col]);
    std::vector<Column> g = gather_columns(ctx, src, fa.sel_out, n_sel);  // one host round trip for all byte totals
    for (size_t u = 0; u < utf8_outs.size(); ++u) {
      g[u].name = out->cols[utf8_outs[u]].name;
      out->cols[utf8_outs[u]] = std::move(g[u]);
    }
  }
  return out;
}

}  // namespace fg

// ================================================================================================
using namespace fg;

extern "C" int flockgpu_filter_project(flockgpu_ctx* ctx, const flockgpu_table* in, const flockgpu_expr* predicate,
                                       const flockgpu_expr* projections, const char* const* out_names, int32_t n_projections,
                                       flockgpu_table** out) {
  return guarded([&] {
    auto c = core_of(ctx);
    FG_CHECK(out, FLOCKGPU_ERR_INVALID, "filter_project: null out pointer");
    FG_CHECK(in && in->table, FLOCKGPU_ERR_INVALID, "filter_project: null input table");
    FG_CHECK(n_projections >= 0 && (n_projections == 0 || projections), FLOCKGPU_ERR_INVALID, "filter_project: bad projection list");
    std::lock_guard<std::recursive_mutex> g(c->mu);
    FG_CUDA(cudaSetDevice(c->device));
    Expr pred;
    if (predicate) pred = tokens_to_expr(predicate);
    std::vector<Expr> projs;
    std::vector<std::string> names;
    for (int i = 0; i < n_projections; ++i) {
      projs.push_back(tokens_to_expr(&projections[i]));
      names.push_back(out_names && out_names[i] ? out_names[i] : "");
    }
    *out = wrap_table(filter_project(c, in->table, predicate ? &pred : nullptr, projs, names));
  });
}
