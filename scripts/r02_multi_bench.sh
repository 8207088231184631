#!/usr/bin/env bash
# bench.py at N GPUs only (no checks).
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps ${2:-3} --warmup ${3:-2} > gpurun_out/r02_bench_c4_n${N}_final.json 2> gpurun_out/r02_bench_c4_n${N}_final.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_c4_n${N}_final.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","phases_ms","clocks","e2e","e2e_nograph"): print(k, json.dumps(d.get(k))[:1500])
except Exception as e:
    print("no bench line", e); print(open("gpurun_out/r02_bench_c4_n${N}_final.err").read()[-3000:])
PY
