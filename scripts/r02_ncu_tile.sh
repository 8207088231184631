#!/usr/bin/env bash
# ncu --set full of the shared-memory tile lag kernel at C4.
set -u
mkdir -p gpurun_out
for rows in ${SC_NCU_ROWS:-1}; do
  SC_BENCH_TILE_ROWS=$rows ncu --set full --clock-control none --import-source on -k regex:lag_tile -s 1 -c 1 \
      -o gpurun_out/r02_lag_tile_rows${rows}_c4 -f python scripts/bench_kernels.py C4 lagtile > gpurun_out/r02_ncu_tile_rows${rows}.log 2>&1
  echo "rows $rows rc=$?"
done
