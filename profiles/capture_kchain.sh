#!/bin/bash
# ncu --set full of the fused chain kernel + its SASS-level page:  bash profiles/capture_kchain.sh <tag>
R=${1:-r02g}
O=gpurun_out; mkdir -p $O
SMALL="python bench.py --rays 2e7 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
$SMALL > $O/${R}_bench_small.json 2> $O/${R}_bench_small.err &&
ncu --set full --clock-control none --import-source on -k regex:k_chain -s 3 -c 1 -o $O/${R}_k_chain -f $SMALL > $O/${R}_ncu_k_chain.log 2>&1
if [ -f $O/${R}_k_chain.ncu-rep ]; then
  python profiles/summarize.py kernel $O/${R}_k_chain.ncu-rep > $O/${R}_k_chain.txt 2>&1
  ncu -i $O/${R}_k_chain.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${R}_k_chain_source.csv.gz
  rm -f $O/${R}_k_chain.ncu-rep
fi
