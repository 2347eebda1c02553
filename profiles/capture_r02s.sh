#!/bin/bash
# r02s: interp with support culling + sort with leaderless ranks: parity and timing
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_interp.py tests/test_gpu_parity.py -m gpu -q -k "interp or griddata or sort or weighted or wavefront" > $O/r02s_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02s_gputest.log
tail -5 $O/r02s_gputest.log
timeout 600 python profiles/interp_probe.py 1e6 512 > $O/r02s_interp_probe.txt 2>&1; cat $O/r02s_interp_probe.txt
timeout 300 python profiles/sort_probe.py 5e7 4 > $O/r02s_sort_probe.txt 2>&1; tail -2 $O/r02s_sort_probe.txt
