#!/bin/bash
# quick GPU iteration: parity suite (optionally a -k filter) + LJ-only bench line
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> [pytest -k expr]'
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then
  (time timeout 600 python -m pytest tests -m gpu -x -q -k "$2") > $out/pytest_gpu.log 2>&1
else
  (time timeout 900 python -m pytest tests -m gpu -x -q) > $out/pytest_gpu.log 2>&1
fi
tail -15 $out/pytest_gpu.log
(time timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-extra) > $out/bench.json 2> $out/bench.err
tail -c 2500 $out/bench.json; tail -3 $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_bench.log 2>&1
