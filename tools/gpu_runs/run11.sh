#!/bin/bash
# run 11 (2 GPUs): NCCL all-to-all + distributed plans parity, weak-scaling bench, sharded q5/q8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_multi.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_multi.txt
tail -15 gpurun_out/pytest_multi.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench11_n2.json 2> gpurun_out/bench11_n2.err ); tail -2 gpurun_out/bench11_n2.err; cat gpurun_out/bench11_n2.json | cut -c1-600
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench11_ref_n2.json 2> gpurun_out/bench11_ref_n2.err ); cat gpurun_out/bench11_ref_n2.json | cut -c1-400
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/nexmark_bench.py --queries q5,q8,q3 --q8-scale 0.125 --reps 10 --no-cpu --no-e2e > gpurun_out/nexmark11_n2.jsonl 2> gpurun_out/nexmark11_n2.err ); tail -3 gpurun_out/nexmark11_n2.err
python - <<PY
import json
for l in open("gpurun_out/nexmark11_n2.jsonl"):
    d=json.loads(l)
    print(d["query"], "n_gpus", d.get("n_gpus"), "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"])
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
