#!/bin/bash
# A/B timing of the fused-chain kernel variants.
#   PXF_CHAIN_VARIANT=<mode><min CTAs/SM><prefetch>   register-file variants (modes 1-3)
#   PXF_CHAIN_VARIANT=4<min CTAs/SM><stages>          bulk-async shared-memory ring (mode 4)
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in ${VARIANTS:-230 432 433 434 442 443 444 446 453}; do
  echo "variant $v"
  PXF_CHAIN_VARIANT=$v python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or chain or config1 or wolter" 2>&1 | tail -1
  PXF_CHAIN_VARIANT=$v $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['trace_kernel_ms'], d['roofline']['frac'], d['config']['hpd'])"
done
