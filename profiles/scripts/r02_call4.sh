cd /root/repo; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_occupancy_gpu.py tests/test_gridprep_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 bash profiles/scripts/run_env.sh MKB_X=1 MKB_X=2 2>&1 | tee gpurun_out/call_env.log
