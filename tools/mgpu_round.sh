#!/bin/bash
# Multi-GPU measurement round (N ranks on one box): communicator probe, peer-optimiser validation, bench lines for the NCCL
# path and the peer-memory optimiser, device timelines (CUPTI) of both.  Output in gpurun_out/mgpu_$N$TAG/
N=${1:-8}
TAG=${2:-}
OUT=gpurun_out/mgpu_$N$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=1
run() {  # name, timeout, env..., -- script args
  name=$1; to=$2; shift 2
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name" | tee -a $OUT/summary.txt
  env "${envs[@]}" timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) "$@" > $OUT/$name.out 2> $OUT/$name.err
  echo "rc=$? $(grep -v '^\[\|NCCL\|^$' $OUT/$name.out | tail -n 1 | cut -c1-700)" | tee -a $OUT/summary.txt
}
run peer_check 200 RB_X=1 -- tools/peer_adam_check.py
run bench_nccl 400 RB_X=1 -- bench.py --gpus $N --steps 300 --warmup 10
run bench_peer 400 RB_X=1 -- bench.py --gpus $N --steps 300 --warmup 10 --peer-optimizer
run timeline_nccl 150 RB_X=1 -- tools/timeline.py --cap 100000 --out $OUT/timeline_nccl.json
run timeline_peer 150 RB_X=1 -- tools/timeline.py --cap 100000 --peer-optimizer --out $OUT/timeline_peer.json
if [ "$N" = "8" ]; then
  run probe_nvls1 150 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,NVLS,ENV NCCL_NVLS_ENABLE=1 -- tools/nccl_probe.py
  run probe_nvls0 150 NCCL_NVLS_ENABLE=0 -- tools/nccl_probe.py
  run bench_nccl_nvls1 400 NCCL_NVLS_ENABLE=1 -- bench.py --gpus $N --steps 300 --warmup 10
fi
cat $OUT/summary.txt | cut -c1-400
