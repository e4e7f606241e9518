# one GPU call: parity of an experimental build (MKB200_LIB) + A/B timing of the in-tree exp_<n>.so builds
set -x
cd /root/repo; mkdir -p gpurun_out
T=${TESTLIB:-F}
export MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_$T.so
timeout 900 python -m pytest tests/test_occupancy_gpu.py tests/test_gridprep_gpu.py -x -q -m gpu > gpurun_out/call_pytest_$T.log 2>&1; echo "pytest $T rc=$?"
tail -5 gpurun_out/call_pytest_$T.log
unset MKB200_LIB
timeout 600 bash profiles/scripts/run_exp.sh > gpurun_out/call_exp.log 2>&1
cat gpurun_out/call_exp.log
