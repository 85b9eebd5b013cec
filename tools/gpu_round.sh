#!/bin/bash
# One GPU visit: tests, smoke, bench (both arms), ncu launch list, ncu --set full of the reciprocal kernels (C2 + C3 meshes)
# and of the cluster pair kernel (dense box).  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
# launch list of two timed steps of one frame each (host-synchronised SCF loop so that every kernel of the WHILE body shows)
ADMP_BENCH_FRAMES=1 ADMP_SCF_HOSTSYNC=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large > $O/launch_plain_$TAG.log 2>&1 && \
ADMP_BENCH_FRAMES=1 ADMP_SCF_HOSTSYNC=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large > $O/ncu_launch_$TAG.log 2>&1
# ncu reports stay on the box (gpurun pulls back at most 64 MiB): export the raw / details pages as text
T=/tmp/ncu_$TAG
mkdir -p $T
export_rep () {   # name
    ncu -i $T/$1.ncu-rep --page raw --csv > $O/$1_$TAG.raw.csv 2>/dev/null
    ncu -i $T/$1.ncu-rep --page details > $O/$1_$TAG.details.txt 2>/dev/null
}
PROF_ONCE=1 python tools/prof_recip.py 1 1 1 1 > $O/prof_c2_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fast_|spread_kernel|gather_kernel' -f -o $T/full_c2 \
    env PROF_ONCE=1 python tools/prof_recip.py 1 1 1 1 > $O/ncu_full_c2_$TAG.log 2>&1
export_rep full_c2
PROF_ONCE=1 python tools/prof_recip.py 2 4 4 1 > $O/prof_c3_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none -k regex:'fast_|spread_kernel|gather_kernel' -f -o $T/full_c3 \
    env PROF_ONCE=1 python tools/prof_recip.py 2 4 4 1 > $O/ncu_full_c3_$TAG.log 2>&1
export_rep full_c3
python tools/pair_roofline.py 32 8.0 > $O/pair_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'pme_cluster_kernel|pme_pair_kernel' -s 20 -c 10 -f -o $T/full_pair \
    python tools/pair_roofline.py 32 8.0 > $O/ncu_full_pair_$TAG.log 2>&1
export_rep full_pair
du -sh $O
tail -3 $O/pytest_$TAG.log
