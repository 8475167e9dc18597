// partition.h -- the multi-way partition pass shared by RepartitionExec on one GPU (partition.cu) and by the
// inter-GPU hash exchange (exchange.cu).
//
// One pass over the relation, whatever the number of partitions (round 1 made one compaction pass and one take() per
// destination: 165 launches on rank 0 for q8 at 8 GPUs):
//   partition_count_kernel    row -> destination byte (the routing function), per-CTA histograms of rows and of the
//                             bytes of every Utf8 column (<= 8 destinations: eight 8-bit counters in one register);
//                             with few destinations (an exchange) the CTA that arrives last also scans
//   partition_scan_kernel     per destination: exclusive scan of the histograms over the CTAs -- a launch of its own
//                             only for many destinations (the 256-way passes of the radix sort)
//   <place>                   decides where this source's rows of every destination go (PartDest): partition.cu lays
//                             the partitions out one behind the other in local buffers, exchange.cu agrees the
//                             layout with the peers and points into their windows
//   partition_scatter_kernel  2048 rows at a time are ordered by destination in shared memory (stable) and written
//                             out segment by segment: a destination receives contiguous, coalesced runs of rows
//                             (fixed-width values; the string bytes staged through shared memory; the Utf8 offsets
//                             through shared memory too, so that consecutive lanes store consecutive entries --
//                             partial-sector stores are what NVLink pays for)
#pragma once

#include "internal.h"
#include "rowkeys.cuh"

namespace fg {

constexpr int PT_THREADS = 256;
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_TILE = 2048;        // rows ordered in shared memory at a time (8 per thread)
constexpr int PT_MAX_PARTS = 256;    // a destination is one byte
constexpr int PT_MAX_UTF8 = 4;       // Utf8 columns one pass can move
constexpr int PT_STAGE_BYTES = 32 * 1024;  // shared staging of a tile's string bytes, twice: source order and destination order (2048 NEXMark strings are ~25 KB)

// Where THIS source's rows of one destination go.
struct PartDest {
  void* val[MAX_IN_COLS];              // fixed-width column c: address of my first row at the destination
  int32_t* off[PT_MAX_UTF8];           // Utf8 column u: the offsets entry of my first row ...
  uint8_t* bytes[PT_MAX_UTF8];         // ... the address of its first byte ...
  long long byte_origin[PT_MAX_UTF8];  // ... and the value of that offsets entry
};

// Geometry and scratch of one pass (host side).
struct PartPass {
  int64_t n_rows = 0;
  int64_t chunk = PT_TILE;  // rows per CTA, a multiple of PT_TILE: CTA b owns rows [b * chunk, (b + 1) * chunk)
  int grid = 1;
  int n_parts = 1;
  std::vector<int> fixed_cols, utf8_cols;  // input columns by kind, in input order
  // columns whose validity bytes travel with the rows: they are moved like 1-byte fixed-width columns, behind the
  // value columns (destination slot fixed_cols.size() + k).  valid_src[k] = the bytes (all ones, synthesised, when an
  // exchange ships the validity of a nullable column that happens to hold no NULL on this rank)
  std::vector<int> valid_cols;
  std::vector<BufferPtr> valid_src;
  BufferPtr pid;       // u8  [n_rows]
  BufferPtr hist;      // u32 [(1 + n_utf8)][grid][n_parts]: rows, then bytes per Utf8 column
  BufferPtr cta_pos;   // u32 same shape: exclusive scan over the CTAs (offset inside my contribution to a destination)
  BufferPtr totals;    // u64 [(1 + n_utf8)][n_parts]
  BufferPtr dest;      // PartDest [n_parts], filled by the place step
};

// Steps 1 and 2.  `dest_rank` >= 0 routes every row to that destination (CoalescePartitionsExec) instead of hashing.
// `ship_nullable`: move validity for every column the SCHEMA calls nullable (an exchange must lay out the same buffers
// on every rank, whatever the data), not only for the columns that hold a NULL here.
PartPass partition_count_scan(const CtxPtr& ctx, const Table& in, const std::vector<int>& routing_cols, int n_parts, int dest_rank = -1, int digit_col = -1,
                              int digit_shift = 0, bool ship_nullable = false);
// One stable radix pass over a relation of fixed-width columns (partition.cu; used by SortExec, sort.cu).
TablePtr radix_pass(const CtxPtr& ctx, const TablePtr& in, int digit_col, int shift);
// Step 4 (after the caller's place step has filled pass.dest).  `abort_flag` (may be NULL): a non-zero word makes the
// kernel return without writing (the exchange sets it when the layout could not be agreed).
void partition_scatter(const CtxPtr& ctx, const Table& in, const PartPass& pass, const unsigned* abort_flag);

}  // namespace fg
