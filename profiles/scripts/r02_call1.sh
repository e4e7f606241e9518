set -x
cd /root/repo; mkdir -p gpurun_out
export MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_C.so
timeout 900 python -m pytest tests/test_occupancy_gpu.py tests/test_gridprep_gpu.py -x -q -m gpu > gpurun_out/call1_pytest_C.log 2>&1; echo "pytest C rc=$?" 
tail -5 gpurun_out/call1_pytest_C.log
unset MKB200_LIB
EXPS="B C D" timeout 600 bash profiles/scripts/run_exp.sh > gpurun_out/call1_exp.log 2>&1
cat gpurun_out/call1_exp.log
