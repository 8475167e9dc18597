// window.cc -- window assembly on the device: hopping / tumbling windows over per-epoch relations that stay in HBM.
//
// The reference assembles a window on the HOST, per invocation: `hopping_window_tasks` keeps a Vec of per-second
// relations, drains `hop_size` of them, appends the new seconds and ships the WHOLE window to the worker function
// again (flock-function/src/aws/window/hopping.rs:54-74; tumbling windows are the hop = size case, tumbling.rs; the
// receiving side collects the pieces of a window in the Arena until its bitmap is full, flock/src/runtime/arena/
// mod.rs:60-85).  With the executor on a GPU the epochs are tables in HBM: a hop uploads only the epochs that are
// new, and the window the plan scans is a device-side concatenation of resident epochs (CoalesceBatchesExec-style
// pointer work plus one D2D copy per column), not a re-feed of `window_size / hop_size` times the data.
#include <deque>

#include "../internal.h"

struct flockgpu_window {
  fg::CtxPtr ctx;
  int window_size = 1, hop_size = 1;  // in epochs (the reference counts seconds)
  std::deque<fg::TablePtr> epochs;    // epochs[0] = first epoch of the next window
  int64_t first_epoch = 0;            // its epoch number
};

using namespace fg;

extern "C" {

int flockgpu_window_open(flockgpu_ctx* ctx, int32_t window_size, int32_t hop_size, flockgpu_window** out) {
  return guarded([&] {
    auto c = core_of(ctx);
    FG_CHECK(out, FLOCKGPU_ERR_INVALID, "window_open: null out pointer");
    FG_CHECK(window_size >= 1 && hop_size >= 1 && hop_size <= window_size, FLOCKGPU_ERR_INVALID,
             "window_open: need 1 <= hop_size <= window_size (hopping.rs:38-43 rejects the rest too), got size %d hop %d", window_size, hop_size);
    auto w = std::make_unique<flockgpu_window>();
    w->ctx = c;
    w->window_size = window_size;
    w->hop_size = hop_size;
    *out = w.release();
  });
}

int flockgpu_window_close(flockgpu_window* w) {
  return guarded([&] {
    if (w && w->ctx) cudaSetDevice(w->ctx->device);
    delete w;
  });
}

int flockgpu_window_push(flockgpu_window* w, flockgpu_table* epoch) {
  return guarded([&] {
    FG_CHECK(w && w->ctx && epoch && epoch->table, FLOCKGPU_ERR_INVALID, "window_push: null argument");
    std::lock_guard<std::recursive_mutex> g(w->ctx->mu);
    if (!w->epochs.empty()) {
      const Table& a = *w->epochs.front();
      const Table& b = *epoch->table;
      FG_CHECK(a.cols.size() == b.cols.size(), FLOCKGPU_ERR_INVALID, "window_push: the epoch has %zu columns, the window %zu", b.cols.size(), a.cols.size());
      for (size_t i = 0; i < a.cols.size(); ++i)
        FG_CHECK(a.cols[i].dtype == b.cols[i].dtype && a.cols[i].name == b.cols[i].name, FLOCKGPU_ERR_INVALID, "window_push: column %zu differs (\"%s\" vs \"%s\")", i,
                 b.cols[i].name.c_str(), a.cols[i].name.c_str());
    }
    w->epochs.push_back(epoch->table);
  });
}

int flockgpu_window_ready(const flockgpu_window* w, int32_t* out) {
  return guarded([&] {
    FG_CHECK(w && out, FLOCKGPU_ERR_INVALID, "window_ready: null argument");
    *out = int(w->epochs.size()) >= w->window_size ? 1 : 0;
  });
}

int flockgpu_window_next(flockgpu_window* w, flockgpu_table** out, int64_t* first_epoch) {
  return guarded([&] {
    FG_CHECK(w && w->ctx && out, FLOCKGPU_ERR_INVALID, "window_next: null argument");
    std::lock_guard<std::recursive_mutex> g(w->ctx->mu);
    FG_CUDA(cudaSetDevice(w->ctx->device));
    FG_CHECK(int(w->epochs.size()) >= w->window_size, FLOCKGPU_ERR_INVALID, "window_next: %zu of %d epochs buffered", w->epochs.size(), w->window_size);
    std::vector<TablePtr> parts(w->epochs.begin(), w->epochs.begin() + w->window_size);
    *out = wrap_table(parts.size() == 1 ? parts[0] : concat_tables(w->ctx, parts));
    if (first_epoch) *first_epoch = w->first_epoch;
    // move the window forward (hopping.rs:60-64: window.drain(..hop_size))
    for (int i = 0; i < w->hop_size; ++i) w->epochs.pop_front();
    w->first_epoch += w->hop_size;
  });
}

}  // extern "C"
