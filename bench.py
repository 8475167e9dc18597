#!/usr/bin/env python
"""bench.py -- NEXMark events/sec through the B200-native executor (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's GPU path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Workload (config.workload = "nexmark_q2_10M_bids"): BASELINE.json configs[1] -- NEXMark q2
(SELECT auction, price FROM bid WHERE auction % 123 = 0) over 10 M bids cut into 64 Ki-row Arrow
batches (152 full + 1 short), one step = one pass of the whole plan over all batches.  With --gpus N
every rank runs the same plan on its own 10 M-bid slice of the stream (RoundRobin sharding, no data-path
collective: SURVEY.md section 8e(i)), so scaling is "weak" and `value` is the sum over ranks.

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  value      bid events/s with the relation already resident in HBM (device leg): K back-to-back executions of the
             plan through ExecutionContext::execute, timed with CUDA events on the library's stream, max over ranks.
             L2 is defeated by rotating over RING distinct resident relations (RING x 80 MB > 126 MB L2).
  e2e        the same metric through the reference-facing call sequence with HOST buffers: feed_data_sources
             (pinned Arrow batches -> HBM), execute, export of the result batch to the host, clean_data_sources.
  roofline   dominant kernel (filter_compact_kernel): algorithmic bytes per launch / its mean launch duration from
             per-launch CUDA events recorded by the library (flockgpu_profile_begin/_end), against the measured HBM
             copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline  the oracle's q2 pipeline on the host cores, same input, same run.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np          # noqa: E402
import pyarrow as pa        # noqa: E402

N_BIDS = 10_000_000          # BASELINE.json configs[1]
BATCH_ROWS = 65536
RING = 4                     # distinct resident relations rotated through (defeats the 126 MB L2)
METRIC = "nexmark_q2_events_per_sec"
UNIT = "events/s"


# Libraries write to fd 1 behind Python's back (NCCL prints "NCCL version ..." there): keep the real stdout for the
# JSON line(s) only and point fd 1 at stderr for everything else.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._stop_evt = threading.Event()
        self.active = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:      # NVML missing: report that honestly
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"} \
            if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else \
                {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                    mask = get(self.h)
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def stop(self) -> dict:
        self._stop_evt.set()
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + getattr(self, "err", "")}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_setup(n_gpus: int):
    """Returns (rank, world, local_rank, dist or None)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0, None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), dist.get_world_size(), local, dist


def dist_max(dist, local: int, value: float) -> float:
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, local: int):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


# ----------------------------------------------------------------------------------------------------------------
def run_gpu(args) -> dict:
    import flock_b200 as fb
    from flock_b200 import nexgen, plans

    rank, world, local, dist = dist_setup(args.gpus)
    ctx = fb.Context(local)
    sampler = ClockSampler(local)
    sampler.start()
    plan = plans.q2()

    # ---- synthetic input: RING distinct 10 M-bid relations per rank (seeded by rank and ring slot)
    t0 = time.time()
    relations = [nexgen.split_batches(nexgen.bids(args.bids, seed=42 + 1000 * rank + r), BATCH_ROWS) for r in range(RING)]
    n_batches = len(relations[0])
    resident = [ctx.import_batches(rel) for rel in relations]
    pinned = [ctx.pinned_copy(b) for b in relations[0]]      # e2e leg: full 4-column bid batches in page-locked host memory
    log(f"[rank {rank}] generated + uploaded {RING} x {args.bids} bids in {time.time() - t0:.1f}s ({n_batches} batches of <= {BATCH_ROWS} rows)")

    ec = fb.ExecutionContext(ctx, plan)

    def step_device(i: int):
        ec.feed_tables([resident[i % RING]])
        return ec.execute_device(0)         # one filter_compact_kernel launch; the survivor count stays in flight

    # survivors per ring slot (untimed): the algorithmic-byte count of SURVEY.md 8(d) needs N_sel
    n_sel = [step_device(r).num_rows for r in range(RING)]
    # ---- device leg: warm-up, then EXACTLY K timed steps between barriers.  Only the last two results are
    # kept alive, as a streaming consumer would: their buffers return to the stream-ordered pool.
    keep = []
    for i in range(args.warmup):
        keep = (keep + [step_device(i)])[-2:]
    ctx.synchronize()
    barrier(dist, local)
    launches0 = ctx.kernel_launches
    sampler.active.set()
    ctx.profile_begin()
    ctx.timer_start(0)
    t_host = time.perf_counter()
    for i in range(args.steps):
        keep = (keep + [step_device(args.warmup + i)])[-2:]
    host_us = (time.perf_counter() - t_host) * 1e6 / args.steps      # host time to ENQUEUE one step (no sync inside)
    ctx.timer_stop(0)
    ctx.synchronize()
    barrier(dist, local)
    sampler.active.clear()
    dev_ms = ctx.timer_ms(0)
    prof = ctx.profile_end()
    launches = ctx.kernel_launches - launches0
    assert keep[-1].num_rows == n_sel[(args.warmup + args.steps - 1) % RING]
    del keep
    dev_ms = dist_max(dist, local, dev_ms)

    # ---- e2e leg: host batches -> HBM -> plan -> host result, every step
    src = [fb.HostRelation(pinned)]      # exported once through the C Data Interface, like FFI structs held by the Rust shim
    # page-locked, uniformly batched host columns stay in host memory: the filter reads `auction` in place over PCIe
    # and fetches `price` only for the survivors (flockgpu_set_option "feed_zero_copy"; --e2e-copy forces the H2D copy)
    ctx.set_option("feed_zero_copy", 0 if args.e2e_copy else 1)
    h2d = sum(b.num_rows for b in pinned) * (8 if args.e2e_copy else 4) + (0 if args.e2e_copy else int(4 * float(np.mean(n_sel))))
    for _ in range(min(args.warmup, 3)):
        ec.feed_data_sources(src)
        res = ec.execute()
        ec.clean_data_sources()
    barrier(dist, local)
    sampler.active.set()
    ctx.profile_begin()
    ctx.timer_start(1)
    t_wall = time.perf_counter()
    d2h = 0
    e2e_steps = args.steps if args.e2e_steps is None else args.e2e_steps
    t_feed = t_exec = t_clean = 0.0
    step_ms = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        ec.feed_data_sources(src)
        t1 = time.perf_counter()
        res = ec.execute()
        t2 = time.perf_counter()
        ec.clean_data_sources()
        t3 = time.perf_counter()
        d2h = sum(b.nbytes for b in res[0])
        t_feed += t1 - t0
        t_exec += t2 - t1
        t_clean += t3 - t2
        step_ms.append((time.perf_counter() - t0) * 1e3)
    ctx.timer_stop(1)
    ctx.synchronize()
    e2e_ms = max(ctx.timer_ms(1), (time.perf_counter() - t_wall) * 1e3)      # host-side work counts too
    e2e_prof = ctx.profile_end()
    barrier(dist, local)
    sampler.active.clear()
    e2e_ms = dist_max(dist, local, e2e_ms)
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel
    peak, peak_src = measured_peak_gbs()
    k = next((v for name, v in prof.items() if name.startswith("filter_compact")), None)
    mean_sel = float(np.mean(n_sel)) if n_sel else 0.0
    alg_bytes = 4.0 * args.bids + 12.0 * mean_sel          # read auction 4 B x N; per survivor read price 4 B, write 8 B
    roofline = None
    traffic = None
    tp = ROOT / "profiles" / "r1_filter_dram_traffic.json"      # dram__bytes_{read,write}.sum of one ncu --set full capture of this kernel
    if tp.exists() and args.bids == N_BIDS:
        try:
            tj = json.loads(tp.read_text())
            traffic = int(tj["dram_bytes_read_per_launch"]) + int(tj["dram_bytes_write_per_launch"])
        except Exception:
            traffic = None
    if k and k["launches"]:
        k_ms = k["ms"] / k["launches"]
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": traffic, "traffic_source": "profiles/r1_filter_dram_traffic.json (ncu --set full, one launch)" if traffic else None,
                    "kernel": "filter_compact_kernel", "kernel_ms": round(k_ms, 5), "launches_timed": k["launches"],
                    "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src}

    result = {
        "metric": METRIC, "value": world * args.bids * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "nexmark_q2_10M_bids" if args.bids == N_BIDS else f"nexmark_q2_{args.bids}_bids", "query": "q2",
                   "bids_per_gpu": args.bids, "batch_rows": BATCH_ROWS, "batches_per_step": n_batches,
                   "sharding": f"round-robin x{world}, no collective", "cache": f"inputs rotate over {RING} resident relations "
                   f"({RING} x {8 * args.bids / 1e6:.0f} MB > 126 MB L2)", "selectivity": mean_sel / args.bids},
        "e2e": {"value": world * args.bids * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "ms_per_step": e2e_ms / max(e2e_steps, 1),
                "host_ms_per_step": {"feed_data_sources": round(t_feed * 1e3 / max(e2e_steps, 1), 4), "execute+export": round(t_exec * 1e3 / max(e2e_steps, 1), 4),
                                     "clean_data_sources": round(t_clean * 1e3 / max(e2e_steps, 1), 4),
                                     "step_median": round(statistics.median(step_ms), 4) if step_ms else None,
                                     "step_max": round(max(step_ms), 4) if step_ms else None},
                "kernels": e2e_prof,
                "feed": "copy: auction + price columns DMA'd to HBM" if args.e2e_copy else
                        "zero-copy: page-locked auction column read in place over PCIe, price fetched for survivors only"},
        "gpu_launches": int(launches), "host_enqueue_us_per_step": round(host_us, 2), "kernels": prof, "clocks": clocks, "roofline": roofline,
        "stream_events_per_sec": world * args.bids * (50 / 46) * args.steps / (dev_ms * 1e-3),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            result["cpu_baseline"] = cpu_baseline(args.bids)
        except Exception as e:                      # the GPU numbers above must still be reported
            log(f"cpu_baseline failed: {e}")
            result["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": f"failed: {e}"[:300]}
    ec.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return result if rank == 0 else None


# ---- the reference's CPU path (oracle port) -----------------------------------------------------------------------
# DataFusion runs one task per partition on a multi-threaded runtime.  The oracle is driven from Python, where a thread
# pool serialises on the interpreter lock (measured: 60 ms per 10 M-bid step on 128 host threads, 150 ms on 8), so the
# timed CPU arm runs the SAME plan JSON with one worker PROCESS per partition: worker p executes the plan over the
# round-robin share of batches the plan's RepartitionExec(RoundRobinBatch(n)) would hand partition p, and the parent
# concatenates the partitions' results like `collect` does.
_REF_BATCHES = None
_REF_EX = None


def _ref_partition(task):
    """Worker: partition `p` of `n` -- feed, execute, clean; returns the result as an Arrow IPC stream."""
    global _REF_EX
    p, n = task
    import oracle
    from flock_b200 import plans
    if _REF_EX is None:
        _REF_EX = oracle.PlanExecutor(plans.q2(1), threads=1)
    share = _REF_BATCHES[p::n] or [_REF_BATCHES[0].slice(0, 0)]
    _REF_EX.feed_data_sources([[share]])
    out = _REF_EX.execute()[0]
    _REF_EX.clean_data_sources()
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, out[0].schema) as w:
        for b in out:
            w.write_batch(b)
    return sink.getvalue().to_pybytes()


def run_reference(args) -> dict | None:
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Rust original cannot be
    built here) on all host cores.  Under torchrun only rank 0 works."""
    global _REF_BATCHES
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import multiprocessing as mp
    import oracle      # builds / loads liboracle.so before the workers are forked
    from flock_b200 import nexgen
    oracle.lib()
    cores = os.cpu_count() or 1
    _REF_BATCHES = nexgen.split_batches(nexgen.bids(args.bids, seed=42), BATCH_ROWS)
    n_parts = min(cores, len(_REF_BATCHES))
    with mp.get_context("fork").Pool(n_parts) as pool:      # forked AFTER the input exists: workers share it copy-on-write
        def step():
            parts = pool.map(_ref_partition, [(p, n_parts) for p in range(n_parts)], chunksize=1)
            return sum(pa.ipc.open_stream(x).read_all().num_rows for x in parts)
        for _ in range(max(args.warmup, 1)):
            step()
        t = time.perf_counter()
        for _ in range(args.steps):
            rows_out = step()
        dt = time.perf_counter() - t
    value = args.bids * args.steps / dt
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": "nexmark_q2_10M_bids" if args.bids == N_BIDS else f"nexmark_q2_{args.bids}_bids", "query": "q2",
                       "bids_per_gpu": args.bids, "batch_rows": BATCH_ROWS, "batches_per_step": len(_REF_BATCHES)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_parts, "kind": "port",
                             "sample": f"all {args.bids} bids per step, {args.steps} steps, one worker process per partition, target_partitions = {n_parts}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "rows_out": rows_out}


def cpu_baseline(n_bids: int, steps: int = 5) -> dict:
    """The CPU arm timed beside the GPU run: `bench.py --impl reference` in a child process (no CUDA context to fork
    around), a bounded sample of the same workload."""
    import subprocess
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", str(steps), "--warmup", "1", "--bids", str(n_bids)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    line = next((l for l in r.stdout.splitlines() if l.startswith("{")), None)
    if r.returncode != 0 or line is None:
        raise RuntimeError(f"reference arm failed (rc={r.returncode}): {r.stderr[-400:]}")
    d = json.loads(line)
    cb = d["cpu_baseline"]
    cb["sample"] += f"; {d['ms_per_step']:.1f} ms per step"
    cb["rows_out"] = d.get("rows_out")
    return cb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--bids", type=int, default=N_BIDS, help="bids per GPU per step (default: the BASELINE.json configuration)")
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the host-buffer leg (default: --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-copy", action="store_true", help="e2e leg: copy the fed columns to HBM instead of reading page-locked batches in place")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gpu" else args.warmup
    res = run_reference(args) if args.impl == "reference" else run_gpu(args)
    if res is not None:
        print(json.dumps(res), file=_JSON_OUT, flush=True)


if __name__ == "__main__":
    main()
