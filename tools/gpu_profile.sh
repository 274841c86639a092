#!/bin/bash
# ncu passes only (never bench values): launch list of graph-replayed updates + --set full of the hand-written kernels.
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches_c2_graph.csv python bench.py --profile-steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_' -c 26 \
    -o $OUT/prof_own python bench.py --profile-steps 1 --profile-mode eager --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -n 6
