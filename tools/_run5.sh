python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2i.log 2>&1; tail -3 gpurun_out/pytest_r2i.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2i_quick.json 2> gpurun_out/bench_r2i_quick.err; echo bench rc=$?
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2i_quick.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'])
for k in ('c1_nonpol_one_gpu','c3_one_gpu','c5_one_gpu','liquid_1024_one_gpu'):
    print(k, str(d.get(k))[:200])
for n,v in d['kernels']['C3'].items(): print(n, v.get('ms'), v.get('frac'))
P
