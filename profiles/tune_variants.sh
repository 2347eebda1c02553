#!/bin/bash
# A/B timing of the fused-chain kernel variants (PXF_CHAIN_VARIANT=<mode><min CTAs/SM><prefetch>)
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in 130 131 141 220 221 230 231 320 321 331; do echo "variant $v"; PXF_CHAIN_VARIANT=$v $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['trace_kernel_ms'], d['roofline']['frac'])"; done
