#!/bin/bash
# ncu --set full of the per-routine kernels the VERDICT asks pipe evidence for (one B200):
#   gpurun --timeout 1500 -- bash profiles/capture_routines.sh r02d "wsprimary:OpWsPrimary ..." [rays]
# Every ncu run is preceded by the same command without ncu (&&), per B200_PROFILING.md.  The .ncu-rep files are
# summarised on the box (raw-page metrics + the gzipped source page) and deleted: seven of them exceed what a
# gpurun call can bring back.
R=${1:-r02d}
LIST=${2:-"wolterprimll:OpLL wsprimary:OpWsPrimary wssecondary:OpWsSecondary tracezern:OpZern woltersine:OpWolterSine refract:OpRefract spocone:OpSpoCone"}
N=${3:-2e7}
O=gpurun_out; mkdir -p $O
for item in $LIST; do
  name=${item%%:*}; op=${item##*:}
  python profiles/routine_probe.py $name 5e7 4 >> $O/${R}_routine_times.txt 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_op<.*${op}[,<>]" -s 1 -c 1 \
      -o $O/${R}_$name -f python profiles/routine_probe.py $name $N 3 > $O/${R}_ncu_$name.log 2>&1
  if [ -f $O/${R}_$name.ncu-rep ]; then
    python profiles/summarize.py kernel $O/${R}_$name.ncu-rep > $O/${R}_k_op_$name.txt 2>&1
    ncu -i $O/${R}_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/${R}_k_op_${name}_source.csv.gz
    rm -f $O/${R}_$name.ncu-rep
  fi
done
cat $O/${R}_routine_times.txt
