#!/bin/bash
# Final single-GPU round of a code freeze: tests, bench lines (C2 headline, C3, C4, reference arm), device timeline, ncu passes.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 $OUT/pytest.log
cp gpurun_out/parity_observed.json $OUT/ 2>/dev/null
timeout 400 python bench.py --steps 300 --warmup 10 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench rc=$?"; grep bench $OUT/bench_c2.err | tail -n 8
timeout 200 python bench.py --impl reference --steps 60 --warmup 3 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err
echo "reference arm rc=$?"
timeout 200 python bench.py --config C3 --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err
timeout 200 python bench.py --config C4 --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err
echo "c3/c4 done"
timeout 200 python tools/timeline.py --out $OUT/timeline_c2.json > $OUT/timeline_c2.txt 2>&1
echo "timeline rc=$?"; grep "^{" $OUT/timeline_c2.txt | cut -c1-300
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches_c2_graph.csv python bench.py --profile-steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_' -c 26 \
    -o $OUT/prof_own python bench.py --profile-steps 1 --profile-mode eager --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -n 8
