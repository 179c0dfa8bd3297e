#!/usr/bin/env bash
# Round-2 GPU call C: scene/abi tests, bench N=1 (both arms, short)
set -u
TAG="${1:-r2c}"; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_scene.py tests/test_gpu_abi.py -m gpu -q > $OUT/${TAG}_pytest_scene.log 2>&1; echo "pytest exit $?"; tail -25 $OUT/${TAG}_pytest_scene.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench_n1.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','kernel_ms','value_public_api','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline'].get('traffic_note'), d.get('reference_cuda',{}).get('speedup_device_path'))"
