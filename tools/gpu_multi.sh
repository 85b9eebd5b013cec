#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): parity of both decompositions, then timings of the big boxes with the x-slab
# scheme and the replicated-mesh atom-block scheme.  Usage: bash tools/gpu_multi.sh TAG N [configs...]
TAG=${1:-rX}; N=${2:-2}; shift; shift
CFGS=${@:-C3 C5}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/run_multigpu.py --check > $O/mg_check_$TAG.log 2>&1; echo "check rc=$?"
tail -5 $O/mg_check_$TAG.log
for c in $CFGS; do
  for s in slab blocks; do
    timeout 900 $TR tools/run_multigpu.py --config $c --scheme $s --steps 2 > $O/mg_${c}_${s}_${N}gpu_$TAG.log 2>&1; echo "$c $s rc=$?"
    tail -2 $O/mg_${c}_${s}_${N}gpu_$TAG.log
  done
done
