#!/bin/bash
# One gpurun call: GPU parity suite, bench line, ncu launch list of the bench command, full captures of the top kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [tests|notests]'
tag=${1:-r01}
mode=${2:-tests}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
if [ "$mode" = tests ]; then
  (time timeout 900 python -m pytest tests -m gpu -x -q) > $out/pytest_gpu.log 2>&1
  tail -5 $out/pytest_gpu.log
fi
(time timeout 600 python bench.py --steps 100 --warmup 20) > $out/bench.json 2> $out/bench.err
tail -c 3000 $out/bench.json
# launch list of the same command (short)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snap > $out/ncu_bench.log 2>&1
# full capture: LJ tile force kernel + tile neighbor build
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lj_tiles|tiles_' -s 4 -c 4 -o $out/lj_full -f \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snap > $out/ncu_lj.log 2>&1
# full capture: SNAP kernels (one step's worth after warm-up)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'snap_(ui|yi|deidrj)_kernel' -s 6 -c 3 -o $out/snap_full -f \
   python tools/snap_time.py 50 50 100 2 > $out/ncu_snap.log 2>&1
ls -la $out
