#!/usr/bin/env bash
# Round-2 evidence call: full GPU suite, both bench arms, parity context numbers, ncu launch list, ncu --set full captures of the
# render kernels on C3 / C4 / C2, compute-sanitizer on small scenes.  Everything lands in gpurun_out/<tag>_*; tools/collect_profiles.py
# turns it into the tracked summaries under profiles/.
set -u
TAG="${1:-r2f}"; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref exit $?"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_c4_n1.json 2> $OUT/${TAG}_bench_c4_n1.err; echo "bench c4 exit $?"
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_c2_n1.json 2> $OUT/${TAG}_bench_c2_n1.err; echo "bench c2 exit $?"
timeout 900 python tools/parity_context.py --out $OUT/${TAG}_parity.json > $OUT/${TAG}_parity.log 2>&1; echo "parity exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
bash tools/gpu_ncu_full.sh c3:64 $OUT/${TAG}_c3_prof auto
bash tools/gpu_ncu_full.sh c4:2 $OUT/${TAG}_c4_prof auto
bash tools/gpu_ncu_full.sh c2:16 $OUT/${TAG}_c2_prof auto
timeout 600 python tools/gpu_latency.py > $OUT/${TAG}_latency.log 2>&1; cp $OUT/small_configs_latency.json $OUT/${TAG}_small_configs_latency.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py > $OUT/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"
GENDR_B200_BWD=ps timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py > $OUT/${TAG}_sanitizer_memcheck_ps.log 2>&1; echo "memcheck(ps) exit $?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_small.py > $OUT/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"
ls -la $OUT | grep ${TAG} | wc -l
