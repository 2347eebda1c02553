#!/bin/bash
# r02m: one-sweep pass under ncu (a middle pass) + launch list of the sort
O=gpurun_out; mkdir -p $O
P="python profiles/sort_probe.py 5e7 2"
timeout 300 $P > $O/r02m_sort_probe.txt 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02m_sort_launches.csv $P > $O/r02m_sort_ncu.log 2>&1
python profiles/summarize.py launches $O/r02m_sort_launches.csv > $O/r02m_sort_launches_summary.txt 2>&1; head -8 $O/r02m_sort_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_onesweep -s 11 -c 1 -o $O/r02m_k_onesweep -f $P > $O/r02m_ncu_k_onesweep.log 2>&1
if [ -f $O/r02m_k_onesweep.ncu-rep ]; then
  python profiles/summarize.py kernel $O/r02m_k_onesweep.ncu-rep > $O/r02m_k_onesweep.txt 2>&1
  ncu -i $O/r02m_k_onesweep.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r02m_k_onesweep_source.csv.gz
  rm -f $O/r02m_k_onesweep.ncu-rep
fi
cat $O/r02m_k_onesweep.txt | head -50
