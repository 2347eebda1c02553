mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_gputest.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err
timeout 600 python bench.py --steps 3 --warmup 3 > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_ncu.log 2>&1
tail -5 gpurun_out/r02a_gputest.log; cat gpurun_out/r02a_bench.json
