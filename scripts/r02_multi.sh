#!/usr/bin/env bash
# Multi-GPU call: equality checks, sharded local Moran / Lee timings, the bench line at N GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -E " ok|Error|error" gpurun_out/r02_dist_check_n$N.log | tail -8
timeout 900 $TR --master-port 29512 scripts/dist_bench.py > gpurun_out/r02_dist_bench_n$N.log 2>&1; echo "dist_bench rc=$?"; tail -n 2 gpurun_out/r02_dist_bench_n$N.log | cut -c1-1200
timeout 1200 $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_c4_n$N.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","phases_ms","e2e","e2e_nograph"): print(k, json.dumps(d.get(k))[:1500])
except Exception as e:
    print("no bench line", e); print(open("gpurun_out/r02_bench_c4_n$N.err").read()[-3000:])
PY
