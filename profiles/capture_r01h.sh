#!/bin/bash
# ncu evidence for the kernels added in the fourth session (run under gpurun on one B200):
#   gpurun --timeout 1500 -- bash profiles/capture_r01h.sh
# Every ncu run is preceded by the same command without ncu (&&), per B200_PROFILING.md.
O=gpurun_out
mkdir -p $O
W="python profiles/whpd_probe.py 5e7 2"
S="python profiles/seg_bench.py 2e7"
Q="python profiles/sort_probe.py 5e7 2"
$W > $O/r01h_whpd.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r01h_whpd_launches.csv $W > $O/r01h_whpd_ncu.log 2>&1
for K in k_wq_collect k_sort_scatter; do
  $W > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o $O/r01h_$K -f $W > $O/r01h_ncu_full_$K.log 2>&1
done
$S > $O/r01h_seg.txt 2>&1
for K in k_chain_seg k_source_seg; do
  $S > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o $O/r01h_$K -f $S > $O/r01h_ncu_full_$K.log 2>&1
done
$Q > $O/r01h_sort.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r01h_sort_launches.csv $Q > $O/r01h_sort_ncu.log 2>&1
