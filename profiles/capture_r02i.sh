#!/bin/bash
# r02i: device set-up sources, pointTo/applyT/indAngle/rmsPoint kernels
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sources or pointto or remaining or vignette" > $O/r02i_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02i_gputest.log
tail -30 $O/r02i_gputest.log
