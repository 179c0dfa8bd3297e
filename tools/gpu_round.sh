#!/usr/bin/env bash
# One gpurun call: GPU parity tests, both bench arms, ncu launch list, one ncu --set full capture of the render kernels.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag> [pytest-args]
set -u
TAG="${1:-run}"; shift || true
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q "$@" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
tail -c 1500 $OUT/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref exit $?"
tail -c 800 $OUT/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -12
