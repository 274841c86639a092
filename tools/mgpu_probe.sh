#!/bin/bash
# Multi-GPU diagnosis run (round 2): communicator bring-up with / without NVLS, peer-memory optimiser validation,
# short bench lines.  Usage (on the GPU box): tools/mgpu_probe.sh N   -> everything lands in gpurun_out/mgpu_N/
N=${1:-4}
OUT=gpurun_out/mgpu_$N
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
  echo "rc=$? $(tail -n 1 $OUT/$name.out | cut -c1-600)" | tee -a $OUT/summary.txt
}
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpus.csv 2>&1
run probe_nvls1 150 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,NVLS,ENV NCCL_NVLS_ENABLE=1 -- tools/nccl_probe.py
run probe_nvls0 150 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV NCCL_NVLS_ENABLE=0 -- tools/nccl_probe.py
run peer_symm 200 RB_PEER_BACKEND=symm -- tools/peer_adam_check.py
run peer_ipc 200 RB_PEER_BACKEND=ipc -- tools/peer_adam_check.py
run bench_nccl 400 RB_X=1 -- bench.py --gpus $N --steps 200 --warmup 5
run bench_peer 400 RB_X=1 -- bench.py --gpus $N --steps 200 --warmup 5 --peer-optimizer
run bench_nccl_nvls1 300 NCCL_NVLS_ENABLE=1 -- bench.py --gpus $N --steps 200 --warmup 5
grep -h "NVLS\|nvls" $OUT/probe_nvls1.err | head -40 > $OUT/nvls_lines.txt
cat $OUT/summary.txt
