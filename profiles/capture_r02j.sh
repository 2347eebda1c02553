#!/bin/bash
# r02j: one-sweep radix sort -- parity, then timing and a launch list
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reconstruct.py -m gpu -q -k "sort or weighted or hpd or reconstruct or southwell" > $O/r02j_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02j_gputest.log
tail -15 $O/r02j_gputest.log
timeout 300 python profiles/sort_probe.py 5e7 4 > $O/r02j_sort_probe.txt 2>&1; cat $O/r02j_sort_probe.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02j_sort_launches.csv python profiles/sort_probe.py 5e7 2 > $O/r02j_sort_ncu.log 2>&1
python profiles/summarize.py $O/r02j_sort_launches.csv > $O/r02j_sort_launches_summary.txt 2>&1; head -12 $O/r02j_sort_launches_summary.txt
