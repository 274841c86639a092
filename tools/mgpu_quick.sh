#!/bin/bash
# Short multi-GPU validation: peer optimiser check (multicast all-gather), default bench line, peer timeline.
N=${1:-2}
OUT=gpurun_out/mgpu_q$N
mkdir -p $OUT
export OMP_NUM_THREADS=1
run() {
  name=$1; to=$2; shift 2
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name" | tee -a $OUT/summary.txt
  env "${envs[@]}" timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) "$@" > $OUT/$name.out 2> $OUT/$name.err
  echo "rc=$? $(grep -v '^\[\|NCCL\|^$' $OUT/$name.out | tail -n 1 | cut -c1-900)" | tee -a $OUT/summary.txt
}
run peer_check 120 RB_X=1 -- tools/peer_adam_check.py
run bench_default 200 RB_X=1 -- bench.py --gpus $N --steps 300 --warmup 10
run timeline_peer 120 RB_X=1 -- tools/timeline.py --cap 100000 --peer-optimizer --out $OUT/timeline_peer.json
tail -n 4 $OUT/peer_check.err | cut -c1-300
