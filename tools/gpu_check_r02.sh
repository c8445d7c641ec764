#!/bin/bash
# full -m gpu suite + the default bench line (with the snap section) on one B200
tag=${1:-r02chk}
out=gpurun_out/$tag; mkdir -p $out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $out/pytest_gpu.log 2>&1
tail -6 $out/pytest_gpu.log
(time timeout 900 python bench.py) > $out/bench.json 2> $out/bench.err; tail -c 3000 $out/bench.json; echo; tail -4 $out/bench.err
