# GPU call: A/B builds + run-time switches + sanitizer on the default build
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_occupancy_gpu.py -x -q -m gpu -k "alternative or nonuniform or compact" 2>&1 | tail -3
EXPS="${EXPS:-R S}" timeout 400 bash profiles/scripts/run_exp.sh 2>&1 | tee gpurun_out/call_exp.log
timeout 200 bash profiles/scripts/run_env.sh MKB_OCC_ZC=2 MKB_OCC_ZC=4 2>&1 | tee -a gpurun_out/call_exp.log
timeout 400 compute-sanitizer --tool memcheck python profiles/scripts/sanitize_occ.py > gpurun_out/san_memcheck_v10.log 2>&1; tail -4 gpurun_out/san_memcheck_v10.log
timeout 400 compute-sanitizer --tool racecheck python profiles/scripts/sanitize_occ.py > gpurun_out/san_racecheck_v10.log 2>&1; tail -4 gpurun_out/san_racecheck_v10.log
