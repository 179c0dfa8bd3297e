#!/usr/bin/env bash
# Round-2 GPU call B: parity of the new kernels + A/B timings (face-stationary vs pixel-stationary backward, build variants)
set -u
TAG="${1:-r2b}"; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not c5_sweep and not axes" > $OUT/${TAG}_pytest_fs.log 2>&1; echo "pytest(fs) exit $?"; tail -4 $OUT/${TAG}_pytest_fs.log
for v in "" "_b3"; do
  for mode in fs ps; do
    GENDR_B200_LIB=$PWD/gendr_b200/libgendr_b200$v.so GENDR_B200_BWD=$mode timeout 300 python tools/gpu_time_kernels.py c3:64 c4:16 c2:16 --tag "lib$v/$mode" 2>&1 | tail -1 | tee -a $OUT/${TAG}_ab.jsonl
  done
done
