#!/usr/bin/env bash
# Round-2 GPU call: tile lag kernels -- correctness per variant, then timing.
set -u
mkdir -p gpurun_out
for pipe in ${SC_PIPES:-1 0}; do
SC_LAG_TILE_PIPE=$pipe timeout 900 python -m pytest tests/test_gpu_lag_tile.py -m gpu -x -q > gpurun_out/r02_pytest_tile_pipe$pipe.log 2>&1; echo "pytest tile pipe=$pipe rc=$?"
tail -n 3 gpurun_out/r02_pytest_tile_pipe$pipe.log
for cfg in ${SC_CFGS:-C4 C2}; do
  SC_BENCH_TILE_ROWS=${SC_ROWS:-0,1} SC_LAG_TILE_PIPE=$pipe timeout 600 python scripts/bench_kernels.py $cfg lagtile > gpurun_out/r02_lagtile_${cfg}_pipe$pipe.json 2> gpurun_out/r02_lagtile_${cfg}_pipe$pipe.err; echo "lagtile $cfg pipe=$pipe rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_lagtile_${cfg}_pipe$pipe.json"))["lag_tiles"]
for k,v in d.items(): print("${cfg} pipe$pipe",k,v)
PY
done
done
