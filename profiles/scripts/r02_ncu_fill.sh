# ncu --set full capture of the default fill kernel on C3 (one launch, after warm-up); LIBX=<exp name> profiles an experimental build
cd /root/repo; mkdir -p gpurun_out
[ -n "$LIBX" ] && export MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_$LIBX.so
timeout 900 ncu --set full --import-source on --clock-control none -k regex:occ_fill_runs -s 4 -c 1 -f -o gpurun_out/fill_${TAG:-x} \
  python bench.py --no-cpu --no-e2e --no-extra --no-scaling --steps 3 --warmup 3 > gpurun_out/ncu_fill_${TAG:-x}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fill_${TAG:-x}.log
