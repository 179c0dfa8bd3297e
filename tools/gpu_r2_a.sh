#!/usr/bin/env bash
# Round-2 GPU call A: full GPU test suite (full-size parity), parity context numbers, C3 + C4 bench lines, ncu capture of C4.
# Usage (repo root on the box): bash tools/gpu_r2_a.sh <tag>
set -u
TAG="${1:-r2a}"
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -30 $OUT/${TAG}_pytest.log
timeout 600 python tools/parity_context.py --out $OUT/${TAG}_parity.json > $OUT/${TAG}_parity.log 2>&1; echo "parity_context exit $?"
tail -12 $OUT/${TAG}_parity.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err; echo "bench c3 exit $?"
python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench_c3.json')); print('C3', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err; echo "bench c4 exit $?"
python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench_c4.json')); print('C4', d['ms_per_step'], d['roofline']['kernel_ms'], d.get('reference_cuda'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 2 -f -o $OUT/${TAG}_c4_prof \
    python bench.py --workload c4 --batch 1 --steps 1 --warmup 3 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_ncu_c4.log 2>&1; echo "ncu c4 exit $?"
ls -la $OUT | tail -12
