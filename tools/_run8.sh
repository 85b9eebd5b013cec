O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu_r2l.json 2> $O/bench_8gpu_r2l.err; echo "bench8 rc=$?"
tail -c 600 $O/bench_8gpu_r2l.json
