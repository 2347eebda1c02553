#!/bin/bash
# r02h: full -m gpu suite, then both bench arms as the driver runs them
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02h_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02h_gputest.log
tail -15 $O/r02h_gputest.log
timeout 600 python bench.py --impl reference > $O/r02h_bench_reference.json 2> $O/r02h_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/r02h_bench.json 2> $O/r02h_bench.err; echo "bench rc=$?"
tail -3 $O/r02h_bench.err
