#!/bin/bash
# r02l: launch-shape sweep (PXF_OP_VARIANT) of the compute-bound per-routine kernels
O=gpurun_out; mkdir -p $O; rm -f $O/r02l_op_variants.txt
for r in refract woltersine spocone conic woltersecondary radgrat; do
  for v in 0 1 2 3 4; do
    echo -n "PXF_OP_VARIANT=$v " >> $O/r02l_op_variants.txt
    PXF_OP_VARIANT=$v timeout 200 python profiles/routine_probe.py $r 5e7 4 >> $O/r02l_op_variants.txt 2>&1
  done
done
cat $O/r02l_op_variants.txt
