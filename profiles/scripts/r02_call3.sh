cd /root/repo; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_occupancy_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 400 bash profiles/scripts/run_env.sh MKB_OCC_CHUNKS=1 MKB_OCC_CHUNKS=2 MKB_OCC_CHUNKS=3 MKB_OCC_CHUNKS=4 2>&1 | tee gpurun_out/call_env.log
