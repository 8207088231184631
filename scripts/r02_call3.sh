#!/usr/bin/env bash
# Round-2 GPU call: pipelined tile kernel profile + the new bench line (all legs).
set -u
mkdir -p gpurun_out
SC_LAG_TILE_PIPE=1 bash scripts/r02_ncu_tile.sh
timeout 1500 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c4.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","phases_ms","roofline","lag_roofline","e2e","e2e_nograph","values_null","lee","nbhd","knn","cpu_baseline","gpu_launches"):
    print(k, json.dumps(d.get(k))[:900])
PY
