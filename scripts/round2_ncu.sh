#!/usr/bin/env bash
# ncu captures for the round-2 lag experiments (run only after round2_first_call.sh has exited 0 without ncu):
#   gpurun --timeout 900 -- 'bash scripts/round2_ncu.sh'
# One `--set full` capture each of: the default lag kernel on 128-byte aligned rows, the row-group lag kernel
# (R = 4) on packed and on aligned rows.  Read the .ncu-rep files here with
#   ncu -i gpurun_out/X.ncu-rep --page details --csv        (summaries go to profiles/)
# and compare l1tex__data_pipe_lsu_wavefronts / l1tex__t_sectors / smsp__warp_issue_stalled_long_scoreboard with
# profiles/r01c_ncu_lag_stat_c4_details.csv (the packed default: 33.7 ms, L1 hit 82 %).
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -c 1"
SC_ROW_ALIGN=32 $NCU -k regex:lag_stat_kernel -s 4 -o gpurun_out/r02_lag_stat_align32_c4 \
    python scripts/bench_kernels.py C4 lag > gpurun_out/r02_ncu_lag_align32.log 2>&1
$NCU -k "regex:lag_group_kernel.*4.*8" -s 2 -o gpurun_out/r02_lag_group4_c4 \
    python scripts/bench_kernels.py C4 laggroup > gpurun_out/r02_ncu_lag_group.log 2>&1
SC_ROW_ALIGN=32 $NCU -k "regex:lag_group_kernel.*4.*8" -s 2 -o gpurun_out/r02_lag_group4_align32_c4 \
    python scripts/bench_kernels.py C4 laggroup > gpurun_out/r02_ncu_lag_group_align32.log 2>&1
ls -la gpurun_out/*.ncu-rep
