# final validation of round 2 on one B200: GPU test suite, smoke(), the bench line, the ncu launch list and one --set full
# capture of the fill kernel (numbers printed under ncu are never bench values)
cd /root/repo; mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final_bench.log | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-scaling > gpurun_out/launches_v10.log 2>&1; echo "ncu list rc=$?"
TAG=v10 bash profiles/scripts/r02_ncu_fill.sh 2>&1 | head -2
