out=gpurun_out/r02last; mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -q -rs) > $out/pytest_gpu.log 2>&1; grep -n "passed\|failed" $out/pytest_gpu.log | tail -2
(timeout 600 python bench.py) > $out/bench.json 2> $out/bench.err; head -c 300 $out/bench.json; echo
(timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline) > $out/bench_20_5.json 2> $out/bench_20_5.err; head -c 300 $out/bench_20_5.json; echo
