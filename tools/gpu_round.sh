#!/bin/bash
# One single-GPU validation round: tensor-core probe (under timeout), full gpu test suite, bench.  Output in gpurun_out/$1/
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 240 python tools/tc_probe.py > $OUT/tc_probe.log 2>&1
PROBE=$?
echo "tc probe rc=$PROBE"; tail -n 25 $OUT/tc_probe.log
if [ $PROBE -ne 0 ]; then export RB_HEAD_TC=0; echo "tensor-core layer 1 DISABLED for the rest of this run"; fi
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 15 $OUT/pytest.log
cp gpurun_out/parity_observed.json $OUT/ 2>/dev/null
timeout 600 python bench.py --steps 300 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; grep bench $OUT/bench.err | tail -n 12
# ---- profiling (never a bench value): launch list of 3 graph-replayed updates, then ncu --set full of the own kernels ----
if [ "${PROFILE:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file $OUT/launches_c2_graph.csv python bench.py --profile-steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_' -c 24 \
      -o $OUT/prof_own python bench.py --profile-steps 1 --profile-mode eager --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ls -la $OUT | tail -n 12
fi
timeout 300 python tools/timeline.py --out $OUT/timeline_c2.json > $OUT/timeline_c2.txt 2>&1
echo "timeline rc=$?"; head -n 3 $OUT/timeline_c2.txt | cut -c1-400
