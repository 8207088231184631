set -u
cd /root/repo
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lee_tc2_kernel -s 2 -c 1 -o gpurun_out/r02_lee_tc2 -f python scripts/_tmp_lee_ncu.py > gpurun_out/r02_lee_tc2_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/r02_lee_tc2_ncu.log
