/* q2_cabi.c -- the C ABI of include/flockgpu.h consumed from C: NEXMark q2 through flock_context_* on hand-built
 * ArrowArrays, exactly what a cgo / Rust `extern "C"` binding would do.  Built and run by
 * tests/test_gpu_cabi.py (`-m gpu`):  gcc -std=c11 -Wall -Werror -Iinclude q2_cabi.c -Lflock_b200 -lflockgpu
 *
 *   usage: q2_cabi <q2_plan.json>          exit code 0 = every check passed
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "flockgpu.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int rc_ = (call);                                                                    \
    if (rc_ != FLOCKGPU_OK) {                                                            \
      fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, flockgpu_last_error()); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

static void no_release(struct ArrowSchema* s) { s->release = NULL; }
static void no_release_array(struct ArrowArray* a) { a->release = NULL; }

enum { N_BATCHES = 3, ROWS = 65536, TAIL = 1234 };

int main(int argc, char** argv) {
  if (argc < 2) return 64;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 65;
  fseek(f, 0, SEEK_END);
  long len = ftell(f);
  fseek(f, 0, SEEK_SET);
  char* plan = (char*)malloc((size_t)len + 1);
  if (fread(plan, 1, (size_t)len, f) != (size_t)len) return 66;
  plan[len] = 0;
  fclose(f);

  /* ---- the `bid` relation (event.rs:336-352): auction, bidder, price Int32; b_date_time Timestamp(ms) */
  static const char* names[4] = {"auction", "bidder", "price", "b_date_time"};
  static const char* formats[4] = {"i", "i", "i", "tsm:"};
  static const char md[] = "\x01\x00\x00\x00\x04\x00\x00\x00name\x03\x00\x00\x00" "bid"; /* {"name": "bid"} */
  struct ArrowSchema fields[4], *field_ptrs[4], schema;
  for (int c = 0; c < 4; ++c) {
    memset(&fields[c], 0, sizeof fields[c]);
    fields[c].format = formats[c];
    fields[c].name = names[c];
    fields[c].release = no_release;
    field_ptrs[c] = &fields[c];
  }
  memset(&schema, 0, sizeof schema);
  schema.format = "+s";
  schema.name = "";
  schema.metadata = md;
  schema.n_children = 4;
  schema.children = field_ptrs;
  schema.release = no_release;

  struct ArrowArray batches[N_BATCHES], cols[N_BATCHES][4], *col_ptrs[N_BATCHES][4];
  const void* col_bufs[N_BATCHES][4][2];
  const void* top_bufs[1] = {NULL};
  const struct ArrowArray* batch_ptrs[N_BATCHES];
  int64_t total = 0, expect = 0;
  int32_t* want_auction = (int32_t*)malloc(sizeof(int32_t) * N_BATCHES * ROWS);
  int32_t* want_price = (int32_t*)malloc(sizeof(int32_t) * N_BATCHES * ROWS);
  uint64_t x = 88172645463325252ull; /* xorshift64 */
  for (int b = 0; b < N_BATCHES; ++b) {
    const int64_t n = b == N_BATCHES - 1 ? TAIL : ROWS;
    int32_t* v32[3];
    int64_t* ts = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    for (int c = 0; c < 3; ++c) v32[c] = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      v32[0][i] = (int32_t)(1000 + (x % 700000)) * ((x >> 40) % 16 == 0 ? -1 : 1);   /* negative ids too: `%` follows the dividend */
      v32[1][i] = (int32_t)(x >> 33);
      v32[2][i] = (int32_t)((x >> 20) % 100000000);
      ts[i] = 1436918400000ll + total + i;
      if ((int64_t)v32[0][i] % 123 == 0) {
        want_auction[expect] = v32[0][i];
        want_price[expect] = v32[2][i];
        ++expect;
      }
    }
    for (int c = 0; c < 4; ++c) {
      memset(&cols[b][c], 0, sizeof cols[b][c]);
      col_bufs[b][c][0] = NULL;
      col_bufs[b][c][1] = c < 3 ? (const void*)v32[c] : (const void*)ts;
      cols[b][c].length = n;
      cols[b][c].n_buffers = 2;
      cols[b][c].buffers = col_bufs[b][c];
      cols[b][c].release = no_release_array;
      col_ptrs[b][c] = &cols[b][c];
    }
    memset(&batches[b], 0, sizeof batches[b]);
    batches[b].length = n;
    batches[b].n_buffers = 1;
    batches[b].buffers = top_bufs;
    batches[b].n_children = 4;
    batches[b].children = col_ptrs[b];
    batches[b].release = no_release_array;
    batch_ptrs[b] = &batches[b];
    total += n;
  }

  /* ---- ExecutionContext: unmarshal -> feed_data_sources -> execute -> clean_data_sources (actor.rs:54-79) */
  flockgpu_ctx* ctx = NULL;
  flock_context* ec = NULL;
  CHECK(flockgpu_open(0, &ctx));
  CHECK(flock_context_unmarshal(ctx, plan, &ec));
  int32_t shuffling = -1;
  CHECK(flock_context_is_shuffling(ec, &shuffling));
  if (shuffling != 0 || flock_context_num_plans(ec) != 1) return 3;
  const struct ArrowSchema* schemas[1] = {&schema};
  const struct ArrowArray* const* sources[1] = {batch_ptrs};
  const int32_t counts[1] = {N_BATCHES};
  const int64_t launches0 = flockgpu_kernel_launches(ctx);
  for (int round = 0; round < 2; ++round) { /* one context serves many invocations */
    CHECK(flock_context_feed_data_sources(ec, schemas, sources, counts, 1));
    flockgpu_table* out = NULL;
    CHECK(flock_context_execute(ec, 0, &out));
    CHECK(flock_context_clean_data_sources(ec));
    if (flockgpu_table_num_rows(out) != expect || flockgpu_table_num_columns(out) != 2) {
      fprintf(stderr, "rows %lld (want %lld), columns %d\n", (long long)flockgpu_table_num_rows(out), (long long)expect, flockgpu_table_num_columns(out));
      return 4;
    }
    struct ArrowSchema os;
    struct ArrowArray oa;
    CHECK(flockgpu_table_export(ctx, out, 0, -1, &os, &oa));
    if (oa.n_children != 2 || strcmp(os.children[0]->name, "auction") || strcmp(os.children[1]->name, "price") || strcmp(os.children[0]->format, "i")) return 5;
    const int32_t* ga = (const int32_t*)oa.children[0]->buffers[1];
    const int32_t* gp = (const int32_t*)oa.children[1]->buffers[1];
    for (int64_t i = 0; i < expect; ++i)
      if (ga[i] != want_auction[i] || gp[i] != want_price[i]) { /* stable: input order, bit-exact */
        fprintf(stderr, "row %lld: (%d, %d), want (%d, %d)\n", (long long)i, ga[i], gp[i], want_auction[i], want_price[i]);
        return 6;
      }
    oa.release(&oa);
    os.release(&os);
    CHECK(flockgpu_table_release(out));
  }
  if (flockgpu_kernel_launches(ctx) - launches0 < 2) return 7; /* the GPU did the work */
  /* an error is a code and a message, never an abort */
  flock_context* bad = NULL;
  if (flock_context_unmarshal(ctx, "{\"execution_plan\": \"cross_join_exec\"}", &bad) != FLOCKGPU_ERR_UNSUPPORTED || !strstr(flockgpu_last_error(), "cross_join_exec")) return 8;
  CHECK(flock_context_free(ec));
  CHECK(flockgpu_close(ctx));
  printf("q2 through the C ABI: %lld of %lld bids kept, bit-exact, %s\n", (long long)expect, (long long)total, flockgpu_version());
  return 0;
}
