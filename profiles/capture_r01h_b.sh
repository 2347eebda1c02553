#!/bin/bash
# second batch of ncu evidence for the fourth session (run under gpurun on one B200):
#   gpurun --timeout 1500 -- bash profiles/capture_r01h_b.sh
O=gpurun_out
mkdir -p $O
W="python profiles/whpd_probe.py 5e7 3"
C="python profiles/config_bench.py 2e7 2 3"
$W > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_wq_collect -s 1 -c 1 -o $O/r01h_k_wq_collect -f $W > $O/r01h_ncu_full_k_wq_collect.log 2>&1
$C > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_program -s 1 -c 1 -o $O/r01h_k_program_zern -f $C > $O/r01h_ncu_full_k_program_zern.log 2>&1
