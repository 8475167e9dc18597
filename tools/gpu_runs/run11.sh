#!/bin/bash
# run 11 (2 GPUs): scanner-CTA prefix; full parity incl. NCCL all-to-all + distributed plans, weak-scaling bench, sharded q5/q8/q3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu11.txt 2>&1 ); echo "pytest rc=$?" >> gpurun_out/pytest_gpu11.txt
tail -12 gpurun_out/pytest_gpu11.txt
( timeout 300 python tools/trace_filter.py gpurun_out/trace11.txt > gpurun_out/trace11_summary.txt 2>&1 ); cat gpurun_out/trace11_summary.txt
( timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 50 > gpurun_out/bench11_n1.json 2> gpurun_out/bench11_n1.err )
( FLOCKGPU_NO_SCANNER=1 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench11_n1_noscan.json 2> gpurun_out/bench11_n1_noscan.err )
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --e2e-steps 50 > gpurun_out/bench11_n2.json 2> gpurun_out/bench11_n2.err ); tail -2 gpurun_out/bench11_n2.err
python - <<PY
import json
for f in ("bench11_n1","bench11_n1_noscan","bench11_n2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "n_gpus", d["n_gpus"], "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],5), "host_us", d.get("host_enqueue_us_per_step"), "roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "e2e ev/s %.3g" % d["e2e"]["value"], d["e2e"].get("host_ms_per_step"), {k:(v["launches"], round(v["ms"]/max(v["launches"],1),4)) for k,v in d["e2e"].get("kernels",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench11_ref_n2.json 2> gpurun_out/bench11_ref_n2.err ); cat gpurun_out/bench11_ref_n2.json | cut -c1-400
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/nexmark_bench.py --queries q5,q8,q3 --q8-scale 0.125 --reps 10 --no-cpu --no-e2e > gpurun_out/nexmark11_n2.jsonl 2> gpurun_out/nexmark11_n2.err ); tail -3 gpurun_out/nexmark11_n2.err
( timeout 900 python tools/nexmark_bench.py --queries q2,q3,q5,q8 --q8-scale 0.125 --reps 10 --no-cpu --no-e2e > gpurun_out/nexmark11_n1.jsonl 2> gpurun_out/nexmark11_n1.err ); tail -3 gpurun_out/nexmark11_n1.err
python - <<PY
import json
for f in ("nexmark11_n1","nexmark11_n2"):
  for l in open("gpurun_out/%s.jsonl"%f):
    d=json.loads(l)
    print(f, d["query"], "n_gpus", d.get("n_gpus"), "ms", round(d["device_ms_median"],4), "rows/s", "%.3g"%d["rows_per_sec"])
    print("   ", {k:(v["launches"], round(v["ms"],4)) for k,v in d["kernels"].items()})
PY
