#!/usr/bin/env bash
# Round-2 GPU call: GPU suite + the bench line with all legs (short run).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_pytest_gpu_full.log
timeout 1500 python bench.py --steps ${SC_STEPS:-1} --warmup ${SC_WARM:-1} > gpurun_out/r02_bench_c4_short.json 2> gpurun_out/r02_bench_c4_short.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_c4_short.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c4_short.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","phases_ms","e2e","values_null","lee","nbhd","knn","local_moran","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:1100])
PY
