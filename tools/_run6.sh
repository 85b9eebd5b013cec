python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'c1',d['c1_nonpol_one_gpu']['evals_per_s'],'c3',d['c3_one_gpu']['ms_per_eval'],'c5',d['c5_one_gpu']['ms_per_eval'],'liq',d['liquid_1024_one_gpu']['evals_per_s'])"
echo "== A/B: PFA Z-inverse with scattered stores"
ADMP_LIB=$PWD/admp_b200/lib_ab/libadmp_b200.so python -m pytest tests/test_gpu_fft.py -x -q 2>&1 | tail -1
for c in "2 4 4" "4 8 8"; do
echo "default $c"; python tools/xpass_time.py $c 2>&1 | grep -E "fft_z_inv|fused"
echo "zinv pfa $c"; ADMP_LIB=$PWD/admp_b200/lib_ab/libadmp_b200.so python tools/xpass_time.py $c 2>&1 | grep -E "fft_z_inv|fused"
done
