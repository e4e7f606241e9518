# one GPU call: parity of experimental builds (TESTLIBS, in-tree exp_<n>.so) + A/B timing (EXPS) through run_exp.sh
set -x
cd /root/repo; mkdir -p gpurun_out
for T in ${TESTLIBS:-F}; do
  MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_$T.so timeout 600 python -m pytest tests/test_occupancy_gpu.py tests/test_gridprep_gpu.py -x -q -m gpu > gpurun_out/call_pytest_$T.log 2>&1; echo "pytest $T rc=$?"
  tail -4 gpurun_out/call_pytest_$T.log
done
timeout 600 bash profiles/scripts/run_exp.sh > gpurun_out/call_exp.log 2>&1
cat gpurun_out/call_exp.log
