for L in 4 6 8; do
ADMP_BENCH_LANES=$L python bench.py --steps 6 --warmup 3 --no-large --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $L value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'c1',d.get('c1_nonpol_one_gpu',{}).get('evals_per_s'))"
done
