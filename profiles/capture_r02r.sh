#!/bin/bash
# r02r: full -m gpu suite, routine table, interp probe, both bench arms
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02r_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02r_gputest.log
tail -8 $O/r02r_gputest.log
timeout 600 python profiles/routine_bench.py 5e7 > $O/r02r_routines.txt 2>&1; tail -30 $O/r02r_routines.txt
timeout 600 python profiles/interp_probe.py 1e6 512 > $O/r02r_interp_probe.txt 2>&1; cat $O/r02r_interp_probe.txt
timeout 600 python bench.py --impl reference > $O/r02r_bench_reference.json 2> $O/r02r_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/r02r_bench.json 2> $O/r02r_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('$O/r02r_bench.json').read().strip().splitlines()[-1]); print('value %.4g ms/step %.3f frac %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])); print({k:(v.get('ms'), v.get('parity',{}).get('ok')) for k,v in d['configs'].items()})"
