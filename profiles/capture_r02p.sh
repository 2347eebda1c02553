#!/bin/bash
# r02p <N>: the host-side ceiling of the N-GPU end-to-end step
N=$1; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR profiles/e2e_probe_multi.py > $O/r02p_e2e_probe_${N}gpu.txt 2> $O/r02p_e2e_probe_${N}gpu.err; echo "rc=$?"
PXF_HOST_THREADS=2 timeout 600 $TR profiles/e2e_probe_multi.py >> $O/r02p_e2e_probe_${N}gpu.txt 2>> $O/r02p_e2e_probe_${N}gpu.err; echo "rc=$?"
grep -v "^\*\*\*\|NCCL" $O/r02p_e2e_probe_${N}gpu.txt; tail -3 $O/r02p_e2e_probe_${N}gpu.err
