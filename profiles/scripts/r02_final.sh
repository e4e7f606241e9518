# final validation of round 2 on one B200: the bench line, one ncu --set full capture of the fill kernel (summarised on the box), the
# ncu launch list, then the GPU test suite and smoke() (numbers printed under ncu are never bench values)
cd /root/repo; mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:occ_fill_runs -s 4 -c 1 -f -o /tmp/fill_v10 python bench.py --no-cpu --no-e2e --no-extra --no-scaling --steps 3 --warmup 3 > gpurun_out/ncu_fill_v10.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py /tmp/fill_v10.ncu-rep gpurun_out/r02_fill_v10 256; cp /tmp/fill_v10.ncu-rep gpurun_out/
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra --no-scaling > gpurun_out/launches_v10.log 2>&1; echo "ncu list rc=$?"
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; head -4 gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
