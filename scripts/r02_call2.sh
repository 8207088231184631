#!/usr/bin/env bash
# Round-2 GPU call: shared-memory tile lag kernel -- correctness, then time per group size / tile tier.
set -u
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for tier in ${SC_TIERS:-0 2}; do
SC_LAG_TILE_TIER=$tier timeout 900 python -m pytest tests/test_gpu_lag_tile.py -m gpu -x -q > gpurun_out/r02_pytest_tile_tier$tier.log 2>&1; echo "pytest tile tier $tier rc=$?"
tail -n 3 gpurun_out/r02_pytest_tile_tier$tier.log
for cfg in ${SC_CFGS:-C4 C2}; do
  SC_BENCH_TILE_ROWS=${SC_ROWS:-0,1} SC_LAG_TILE_TIER=$tier timeout 600 python scripts/bench_kernels.py $cfg lagtile > gpurun_out/r02_lagtile_${cfg}_tier$tier.json 2> gpurun_out/r02_lagtile_${cfg}_tier$tier.err; echo "lagtile $cfg tier $tier rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_lagtile_${cfg}_tier$tier.json"))["lag_tiles"]
for k,v in d.items(): print("${cfg} tier$tier",k,v)
PY
done
done
