#!/bin/bash
# full ncu capture of the SNAP kernels (regex) at 250 000 atoms + source page:  bash tools/gpu_ncu_snap.sh <tag> <regex> [skip] [count]
tag=${1:-ncusnap}; rx=${2:-snap_yi}; skip=${3:-0}; cnt=${4:-1}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o $out/one -f \
   python tools/snap_time.py 50 50 100 1 > $out/ncu.log 2>&1
ncu -i $out/one.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
ncu -i $out/one.ncu-rep --page source --csv > $out/source.csv 2>/dev/null
python tools/ncu_summary.py $out/raw.csv > $out/summary.csv
rm -f $out/one.ncu-rep
ls -la $out
