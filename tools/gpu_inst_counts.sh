#!/usr/bin/env bash
# instruction counts / lanes / duration of the render kernels on a workload, per library variant and backward mode (ncu, 3 metrics)
# usage: tools/gpu_inst_counts.sh <workload:batch> <outfile> [lib suffixes...]
WL="$1"; OUT="$2"; shift 2
for v in "$@"; do
  [ "$v" = "main" ] && v=""
  for mode in fs ps; do
    GENDR_B200_LIB=$PWD/gendr_b200/libgendr_b200$v.so GENDR_B200_BWD=$mode timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:render -s 4 -c 2 --csv python tools/gpu_time_kernels.py $WL --n 1 2>/dev/null | grep -E "render" | \
      python -c "
import sys,csv
rows=list(csv.reader(sys.stdin))
for r in rows: print('lib$v/$mode', r[4][:60], r[-3], r[-1])
" | tee -a $OUT
  done
done
