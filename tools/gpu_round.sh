#!/bin/bash
# One GPU visit: tests, bench (both arms), ncu launch list, ncu --set full of the reciprocal kernels (C2 + C3 meshes)
# and of the pair kernel (dense box).  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_$TAG.log
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
python tools/pair_roofline.py 32 8.0 > $O/pair_plain_$TAG.log 2>&1
ADMP_SCF_HOSTSYNC=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-large > $O/ncu_launch_$TAG.log 2>&1
# ncu reports stay on the box (gpurun pulls back at most 64 MiB): export the raw / details / source pages as text
T=/tmp/ncu_$TAG
mkdir -p $T
export_rep () {   # name
    ncu -i $T/$1.ncu-rep --page raw --csv > $O/$1_$TAG.raw.csv 2>/dev/null
    ncu -i $T/$1.ncu-rep --page details > $O/$1_$TAG.details.txt 2>/dev/null
}
ncu --set full --clock-control none --import-source on -k regex:'fast_|spread_kernel|gather_kernel' -f -o $T/full_c2 \
    env PROF_ONCE=1 python tools/prof_recip.py 1 1 1 1 > $O/ncu_full_c2_$TAG.log 2>&1
export_rep full_c2
ncu -i $T/full_c2.ncu-rep --page source --csv --kernel-name regex:fast_x_conv > $O/full_c2_xconv_$TAG.source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'fast_|spread_kernel|gather_kernel' -f -o $T/full_c3 \
    env PROF_ONCE=1 python tools/prof_recip.py 2 4 4 1 > $O/ncu_full_c3_$TAG.log 2>&1
export_rep full_c3
ncu --set full --clock-control none --import-source on -k regex:pme_pair_kernel -c 2 -f -o $T/full_pair \
    python tools/pair_roofline.py 32 8.0 > $O/ncu_full_pair_$TAG.log 2>&1
export_rep full_pair
ncu -i $T/full_pair.ncu-rep --page source --csv --kernel-name regex:pme_pair_kernel --launch-count 1 > $O/full_pair_$TAG.source.csv 2>/dev/null
du -sh $O
tail -3 $O/pytest_$TAG.log
