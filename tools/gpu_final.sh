#!/usr/bin/env bash
# Full round check on the GPU box: smoke, the whole -m gpu suite, both bench arms, launch list, ncu full, sanitizer on the new kernels.
set -u
TAG="${1:-final}"; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"; cut -c1-400 $OUT/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref exit $?"; cut -c1-300 $OUT/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_ncu_full.log 2>&1
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python -m pytest -m gpu -q -x tests/test_gpu_voxel.py::test_voxelizer_bit_identical_to_reference_golden \
      "tests/test_gpu_scene.py::test_fused_scene_equals_staged_kernels" tests/test_gpu_scene.py::test_camera_and_lighting_kernels_vs_reference_golden \
      > $OUT/${TAG}_sanitizer_${tool}.log 2>&1; echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_sanitizer_${tool}.log | tail -2
done
