#!/bin/bash
# r02b: full GPU suite (no -x), per-routine table, sort probe
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r02b_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_gputest.log
timeout 600 python profiles/routine_bench.py 5e7 > $O/r02b_routines.txt 2>&1
timeout 300 python profiles/sort_probe.py 5e7 2 > $O/r02b_sort.txt 2>&1
tail -8 $O/r02b_gputest.log
