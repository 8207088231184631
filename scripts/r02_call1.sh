#!/usr/bin/env bash
# Round-2 GPU call 1: the lag-kernel experiments written at the end of round 1 (row alignment, row groups).
set -u
mkdir -p gpurun_out
SC_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_zz_gpu_experimental.py -m gpu -q > gpurun_out/r02_pytest_experimental.log 2>&1; echo "pytest experimental rc=$?"
for cfg in C4 C2; do
  timeout 600 python scripts/bench_kernels.py $cfg lag,values,rows > gpurun_out/r02_kernels_${cfg}_align8.json 2> gpurun_out/r02_kernels_${cfg}_align8.err
  SC_ROW_ALIGN=32 timeout 600 python scripts/bench_kernels.py $cfg lag,values,rows,lagsweep > gpurun_out/r02_kernels_${cfg}_align32.json 2> gpurun_out/r02_kernels_${cfg}_align32.err
  timeout 600 python scripts/bench_kernels.py $cfg laggroup > gpurun_out/r02_laggroup_${cfg}_align8.json 2> gpurun_out/r02_laggroup_${cfg}_align8.err
  SC_ROW_ALIGN=32 timeout 600 python scripts/bench_kernels.py $cfg laggroup > gpurun_out/r02_laggroup_${cfg}_align32.json 2> gpurun_out/r02_laggroup_${cfg}_align32.err
done
SC_ORDER_POINTS_PER_CELL=2 timeout 600 python scripts/bench_kernels.py C4 laggroup > gpurun_out/r02_laggroup_C4_ppc2.json 2> gpurun_out/r02_laggroup_C4_ppc2.err
SC_ROW_ALIGN=32 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_align32.log 2>&1; echo "pytest align32 rc=$?"
tail -3 gpurun_out/r02_pytest_experimental.log gpurun_out/r02_pytest_gpu_align32.log
