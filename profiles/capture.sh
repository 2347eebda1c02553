#!/bin/bash
# Reproduces the profiles in this directory (run under gpurun on one B200):
#   gpurun --timeout 1500 -- bash profiles/capture.sh r01g
# Every ncu run is preceded by the same command without ncu (&&), per B200_PROFILING.md.
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
SMALL="python bench.py --rays 2e7 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
FULL="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$SMALL > $O/${R}_bench_small.json 2> $O/${R}_bench_small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${R}_launches.csv $SMALL > $O/${R}_ncu_launches.log 2>&1
# the same launch list at the bench's own size (1.25e8 rays/launch), warm L2 between kernels
$FULL > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
    --log-file $O/${R}_launches_full.csv $FULL > $O/${R}_ncu_launches_full.log 2>&1
for K in k_chain k_bracket_collect k_small_select k_cand_hist k_cand_finish; do
  $SMALL > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 \
      -o $O/${R}_$K -f $SMALL > $O/${R}_ncu_full_$K.log 2>&1
done
