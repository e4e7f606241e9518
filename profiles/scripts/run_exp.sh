for n in ${EXPS:-A B C D}; do
  MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_$n.so python bench.py --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', d['roofline']['kernel_ms_mean'], d['roofline']['frac'])"
done
