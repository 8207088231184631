#!/usr/bin/env bash
# Round-2 evidence run on one B200: the default bench line, then (only after it exited 0) the ncu launch list of the
# same command and one `--set full` capture each of the lag tile kernel and the CTA-pair Lee kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/r02_bench_c4_1gpu.err; rc=$?; echo "bench rc=$rc"
[ $rc -eq 0 ] || { tail -20 gpurun_out/r02_bench_c4_1gpu.err; exit 1; }
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_c4_reference.json 2> gpurun_out/r02_bench_c4_reference.err; echo "reference rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c4_default.csv \
    python bench.py --no-cpu > gpurun_out/r02_ncu_bench.out 2> gpurun_out/r02_ncu_bench.err; echo "ncu launch list rc=$?"
SC_BENCH_TILE_ROWS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lag_tile_kernel -s 3 -c 1 \
    -o gpurun_out/r02_lag_tile_c4 -f python scripts/bench_kernels.py C4 lagtile > gpurun_out/r02_ncu_lag_tile.log 2>&1; echo "ncu lag rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lee_tc2_kernel -s 2 -c 1 \
    -o gpurun_out/r02_lee_tc2_c3 -f python scripts/lee_tc_prof.py > gpurun_out/r02_ncu_lee_tc2.log 2>&1; echo "ncu lee rc=$?"
timeout 600 python scripts/bench_kernels.py C4 lagtile,values,rows > gpurun_out/r02_kernels_c4.json 2> gpurun_out/r02_kernels_c4.err; echo "kernels C4 rc=$?"
timeout 600 python scripts/bench_kernels.py C2 lagtile,values,rows > gpurun_out/r02_kernels_c2.json 2> gpurun_out/r02_kernels_c2.err; echo "kernels C2 rc=$?"
timeout 600 python bench.py --workload C2 --no-legs > gpurun_out/r02_bench_c2_1gpu.json 2> gpurun_out/r02_bench_c2_1gpu.err; echo "bench C2 rc=$?"
