#!/usr/bin/env bash
# one ncu --set full capture of the render kernels of a workload: tools/gpu_ncu_full.sh <workload:batch> <out-prefix> [fs|ps]
WL="$1"; OUT="$2"; MODE="${3:-fs}"
GENDR_B200_BWD=$( [ "$MODE" = auto ] && echo "" || echo $MODE ) timeout 900 ncu --set full --clock-control none --import-source on -k regex:render -s 4 -c 2 -f -o $OUT \
   python tools/gpu_time_kernels.py $WL --n 1 > ${OUT}.log 2>&1
echo "ncu exit $?"
