# A/B runs of run-time switches (environment variables) on the default library; BATCH=<pockets> for small batches
for e in "$@"; do
  env $e python bench.py --no-cpu --no-e2e --no-extra --no-scaling --steps 20 ${BATCH:+--batch $BATCH} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e', d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['prep_ms_mean'], d['roofline']['frac'])"
done
