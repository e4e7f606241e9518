cd /root/repo; mkdir -p gpurun_out
export MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_X.so
timeout 100 python -m pytest tests/test_distance_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 60 python profiles/scripts/k3_time.py 2>&1 | tail -2 | tee gpurun_out/k3_x2.log
