#!/bin/bash
# r02v: end-of-round evidence -- full -m gpu suite, smoke, routine table, both bench arms, sort launch list + k_onesweep ncu
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02v_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02v_gputest.log
tail -6 $O/r02v_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02v_smoke.log 2>&1; tail -2 $O/r02v_smoke.log
timeout 600 python profiles/routine_bench.py 5e7 > $O/r02v_routines.txt 2>&1; tail -3 $O/r02v_routines.txt
timeout 600 python bench.py --impl reference > $O/r02v_bench_reference.json 2> $O/r02v_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/r02v_bench.json 2> $O/r02v_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('$O/r02v_bench.json').read().strip().splitlines()[-1]); print('value %.4g ms/step %.3f frac %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])); print({k:(round(v.get('ms'),2), v.get('parity',{}).get('ok')) for k,v in d['configs'].items()})"
P="python profiles/sort_probe.py 5e7 2"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02v_sort_launches.csv $P > $O/r02v_sort_ncu.log 2>&1
python profiles/summarize.py launches $O/r02v_sort_launches.csv > $O/r02v_sort_launches_summary.txt 2>&1; head -5 $O/r02v_sort_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_onesweep -s 11 -c 1 -o $O/r02v_k_onesweep -f $P > $O/r02v_ncu_k_onesweep.log 2>&1
if [ -f $O/r02v_k_onesweep.ncu-rep ]; then
  python profiles/summarize.py kernel $O/r02v_k_onesweep.ncu-rep > $O/r02v_k_onesweep.txt 2>&1
  rm -f $O/r02v_k_onesweep.ncu-rep
fi
head -22 $O/r02v_k_onesweep.txt
