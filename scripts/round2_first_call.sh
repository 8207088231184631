#!/usr/bin/env bash
# First GPU call of the next round (everything below was written after round 1's GPU budget was spent):
#   gpurun --timeout 2700 -- 'bash scripts/round2_first_call.sh'      (about 25-35 box minutes)
# 1. the GPU suite and the default bench (now with the values_null leg);
# 2. the row-alignment experiment (SC_ROW_ALIGN=32: 128-byte aligned rows) on the lag kernel, the
#    value-permuting null and the graph-row null, C4 and C2, plus the parity tests under that alignment;
# 3. the row-group lag kernel (SC_LAG_GROUP=2|4|8): its gated tests, then time per group size;
# Multi-GPU follow-up (a second call with --gpus 2):
#   gpurun --gpus 2 --timeout 600 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#       --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r02_dist_check.log 2>&1'
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; echo "bench rc=$?"
for cfg in C4 C2; do
  python scripts/bench_kernels.py $cfg lag,values,rows > gpurun_out/r02_kernels_${cfg}_align8.json 2> gpurun_out/r02_kernels_${cfg}_align8.err
  SC_ROW_ALIGN=32 python scripts/bench_kernels.py $cfg lag,values,rows,lagsweep > gpurun_out/r02_kernels_${cfg}_align32.json 2> gpurun_out/r02_kernels_${cfg}_align32.err
done
SC_TEST_EXPERIMENTAL=1 python -m pytest tests/test_zz_gpu_experimental.py -m gpu -q > gpurun_out/r02_pytest_experimental.log 2>&1; echo "pytest experimental rc=$?"
for cfg in C4 C2; do
  python scripts/bench_kernels.py $cfg laggroup > gpurun_out/r02_laggroup_${cfg}_align8.json 2> gpurun_out/r02_laggroup_${cfg}_align8.err
  SC_ROW_ALIGN=32 python scripts/bench_kernels.py $cfg laggroup > gpurun_out/r02_laggroup_${cfg}_align32.json 2> gpurun_out/r02_laggroup_${cfg}_align32.err
  SC_ORDER_POINTS_PER_CELL=2 python scripts/bench_kernels.py $cfg laggroup > gpurun_out/r02_laggroup_${cfg}_ppc2.json 2> gpurun_out/r02_laggroup_${cfg}_ppc2.err
done
SC_ROW_ALIGN=32 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r02_pytest_gpu_align32.log 2>&1; echo "pytest align32 rc=$?"
tail -3 gpurun_out/r02_pytest_gpu.log gpurun_out/r02_pytest_experimental.log gpurun_out/r02_pytest_gpu_align32.log
