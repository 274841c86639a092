#!/bin/bash
# compute-sanitizer memcheck over the small-size parity tests of the replay kernels, the C51 loss, the
# optimiser and the fused head (forward incl. the tcgen05 layer 1, backward).  Never a timing run.
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTORCH_NO_CUDA_MEMORY_CACHING=1        # every tensor its own cudaMalloc: out-of-bounds accesses are not hidden by the caching allocator
T=tests/test_gpu_parity.py
timeout ${LIMIT:-70} compute-sanitizer --tool memcheck --error-exitcode 86 --log-file $OUT/memcheck.log \
  python -u -m pytest -v -p no:cacheprovider -m gpu \
  "$T::test_tree_update_find_golden" "$T::test_replay_sample_golden" "$T::test_append_golden" "$T::test_c51_golden" \
  "$T::test_clip_adam_oracle" "$T::test_fused_head_backward" > $OUT/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"
grep -E "PASSED|FAILED|ERROR|passed|failed" $OUT/memcheck_pytest.log | tail -n 30
tail -n 5 $OUT/memcheck.log
