#!/usr/bin/env bash
# A/B timing of alternative builds of the library on the bench workloads: tools/gpu_ab.sh <lib1> <lib2> ...  (paths relative to repo root)
for lib in "$@"; do
  for wl in c3 c2; do
    GENDR_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-reference-cuda 2>&1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', '$wl', 'ms/step %.3f' % d['ms_per_step'], 'bwd %.3f' % d['roofline']['kernel_ms'], 'e2e %.3f' % d['e2e']['ms_per_step'])"
  done
done
