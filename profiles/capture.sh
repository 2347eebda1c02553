#!/bin/bash
# Reproduces the profiles in this directory (run under gpurun on one B200):
#   gpurun --timeout 1500 -- bash profiles/capture.sh r01
# Every ncu run is preceded by the same command without ncu (&&), per B200_PROFILING.md.
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
SMALL="python bench.py --rays 2e7 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$SMALL > $O/${R}_bench_small.json 2> $O/${R}_bench_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${R}_launches.csv $SMALL > $O/${R}_ncu_launches.log 2>&1
$SMALL > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_chain -s 3 -c 1 \
    -o $O/${R}_k_chain -f $SMALL > $O/${R}_ncu_full.log 2>&1
$SMALL > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bracket_collect -s 3 -c 1 \
    -o $O/${R}_k_bracket_collect -f $SMALL > $O/${R}_ncu_full_select.log 2>&1
