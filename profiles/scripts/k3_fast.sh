python -m pytest tests/test_distance_gpu.py -m gpu -x -q 2>&1 | tail -8
python bench.py --no-cpu --no-e2e --no-scaling --steps 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ex=d.get('extra') or d.get('config',{}).get('extra'); 
for k,v in (ex or {}).items():
    if k.startswith('c4a'): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
"
