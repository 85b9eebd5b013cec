python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for q in 1 0; do
ADMP_FFT_QUICKVIR=$q python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('quickvir $q value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'c1',d['c1_nonpol_one_gpu']['evals_per_s'],'c3',d['c3_one_gpu']['ms_per_eval'],'c5',d['c5_one_gpu']['ms_per_eval'],'liq',d['liquid_1024_one_gpu']['evals_per_s'], 'E', d['config']['energy_last_frame'], d['c3_one_gpu']['energy_per_replica'])"
done
