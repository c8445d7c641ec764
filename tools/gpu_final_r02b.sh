#!/bin/bash
# final round-2 evidence on one B200: full -m gpu suite, default bench line (with the snap / weak16M / snap_strong sections), the
# driver-style 20/5 window, the reference arm, smoke()
tag=${1:-r02final}
out=gpurun_out/$tag; mkdir -p $out
(time timeout 1200 python -m pytest tests -m gpu -q -rs) > $out/pytest_gpu.log 2>&1
tail -4 $out/pytest_gpu.log
(time timeout 900 python bench.py) > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json; echo; tail -3 $out/bench.err
(time timeout 600 python bench.py --steps 20 --warmup 5) > $out/bench_20_5.json 2> $out/bench_20_5.err; head -c 400 $out/bench_20_5.json; echo
(time timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > $out/bench_reference.json 2> $out/bench_reference.err; head -c 400 $out/bench_reference.json; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
