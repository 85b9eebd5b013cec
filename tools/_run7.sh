O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/run_multigpu.py --check > $O/mg_check_r2k.log 2>&1; echo "check rc=$?"
tail -4 $O/mg_check_r2k.log
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu_r2k.json 2> $O/bench_2gpu_r2k.err; echo "bench2 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_2gpu_r2k.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'])
for k in ('c3_slab','c5_slab'):
    print(k, {a:b for a,b in d[k].items() if a in ('ms_per_eval','n_gpus','speedup_vs_one_gpu','stage_ms_rank0')})
P
