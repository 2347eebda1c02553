#!/bin/bash
# r02f: Newton loops without per-division branches -- parity, then the headline kernel time
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x -k "not config and not reconstruct and not dist" > $O/r02f_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02f_gputest.log
tail -5 $O/r02f_gputest.log
python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > $O/r02f_bench.json 2> $O/r02f_bench.err
python -c "
import json; d=json.load(open('$O/r02f_bench.json')); print('trace_kernel_ms', d['config']['trace_kernel_ms'], 'ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['kernel'][:60])"
python profiles/routine_bench.py 5e7 > $O/r02f_routines.txt 2>&1; head -12 $O/r02f_routines.txt
