for n in ${EXPS:-A B C D}; do
  MKB200_LIB=/root/repo/moleculekit_b200/lib/exp_$n.so python bench.py --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['extra']; print('$n', 'dist', round(e['c4a_distances']['kernel_ms'],4), 'contacts', round(e['c4a_contacts']['kernel_ms'],4), 'sparse', round(e['c4b_sparse_contacts']['ms_per_step'],3))"
done
