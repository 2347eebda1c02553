#!/bin/bash
# r02q: interpolateVec / wavefront parity
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_interp.py tests/test_gpu_reconstruct.py -m gpu -q > $O/r02q_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02q_gputest.log
tail -60 $O/r02q_gputest.log
