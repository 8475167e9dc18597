#!/bin/bash
# run 15 (8 GPUs): BASELINE configs[4] -- q8 over 1 B events sharded by hash radix across 8 x B200 (NCCL all-to-all), plus q5 / q3 sharded
# and the q2 weak-scaling bench line at N = 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/nexmark_bench.py --queries q8,q5,q3 --q8-scale 1.0 --reps 5 --no-cpu --no-e2e > gpurun_out/nexmark15_n8.jsonl 2> gpurun_out/nexmark15_n8.err ); tail -4 gpurun_out/nexmark15_n8.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/nexmark15_n8.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["query"], "n_gpus", d.get("n_gpus"), "scale", d["scale"], "rows_in", d["rows_in"], "rows_out", d["rows_out"], "ms", round(d["device_ms_median"],4), "best", round(d["device_ms_best"],4), "rows/s", "%.3g"%d["rows_per_sec"], "events/s", "%.3g"%d["stream_events_per_sec"])
    print("   kernel sum (rank 0) ms", round(sum(v["ms"] for v in d["kernels"].values()),4), {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 200 --warmup 5 --e2e-steps 20 > gpurun_out/bench15_n8.json 2> gpurun_out/bench15_n8.err ); tail -2 gpurun_out/bench15_n8.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/bench15_n8.json"))
print("n_gpus", d["n_gpus"], "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],5), "roofline", d["roofline"]["frac"], "e2e ev/s %.3g" % d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
