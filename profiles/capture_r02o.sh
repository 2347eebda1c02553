#!/bin/bash
# r02o <N>: N-GPU run -- NCCL parity test (N>=2), then both bench arms exactly as the driver launches them
N=$1; O=gpurun_out; mkdir -p $O
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q > $O/r02o_gputest_dist.log 2>&1; echo "pytest rc=$?" >> $O/r02o_gputest_dist.log
  tail -4 $O/r02o_gputest_dist.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus $N --impl reference --steps 20 --warmup 3 > $O/r02o_bench_reference_${N}gpu.json 2> $O/r02o_bench_reference_${N}gpu.err; echo "ref rc=$?"
timeout 900 $TR bench.py --gpus $N --no-configs --no-cpu-baseline > $O/r02o_bench_${N}gpu.json 2> $O/r02o_bench_${N}gpu.err; echo "bench rc=$?"
tail -2 $O/r02o_bench_${N}gpu.err
python - <<PY
import json
for f in ("$O/r02o_bench_reference_${N}gpu.json", "$O/r02o_bench_${N}gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms/step", d.get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"), "cores", d.get("cpu_baseline", {}).get("cores"), d.get("parity_check"))
    except Exception as e:
        print(f, "unreadable", e)
PY
