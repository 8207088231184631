set -u
cd /root/repo
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_smoke.log
bash scripts/r02_final_1gpu.sh
