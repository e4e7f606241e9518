# A/B runs of alternative in-tree builds (moleculekit_b200/lib/exp_<n>.so), selected at run time with MKB200_LIB
python bench.py --no-cpu --no-e2e --no-extra --no-scaling --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'])"
for n in ${EXPS:-A B C D}; do
  MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_$n.so python bench.py --no-cpu --no-e2e --no-extra --no-scaling --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'])"
done
