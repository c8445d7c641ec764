#!/bin/bash
# full ncu capture of the tile kernels during a short bench run:  gpurun -- 'bash tools/gpu_ncu_tiles.sh <tag> [skip] [count]'
tag=${1:-ncu}; skip=${2:-0}; cnt=${3:-8}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lj_tiles|tiles_(search|lists|fill)' -s $skip -c $cnt -o $out/tiles_full -f \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu.log 2>&1
ncu -i $out/tiles_full.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
python tools/ncu_summary.py $out/raw.csv > $out/summary.csv
ls -la $out
