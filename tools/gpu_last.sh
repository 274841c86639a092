#!/bin/bash
# Shortest useful single-GPU check of a small kernel change: gpu tests, one headline bench line, device timeline.
TAG=${1:-last}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 150 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/pytest.log
timeout 100 python bench.py --steps 300 --warmup 10 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench rc=$?"; cut -c1-400 $OUT/bench_c2.json
timeout 60 python tools/timeline.py --out $OUT/timeline_c2.json > $OUT/timeline_c2.txt 2>&1
echo "timeline rc=$?"; grep -E "conv_wgrad|^\{" $OUT/timeline_c2.txt | cut -c1-200
