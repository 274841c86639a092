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
