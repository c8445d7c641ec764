# end-of-round evidence: parity suite, the default bench line (with SNAP + CPU baselines), reference arm, launch list,
# and one large-brick run (configs[3] per-GPU brick: 160^3 cells = 16 384 000 atoms)
tag=${1:-r01z}
out=gpurun_out/$tag; mkdir -p $out
(time timeout 600 python -m pytest tests -m gpu -x -q) > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
(time timeout 600 python bench.py) > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json; echo
(time timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 400 $out/bench_reference.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snap > $out/ncu_bench.log 2>&1
(time timeout 400 python bench.py --region 160 160 160 --steps 40 --warmup 20 --no-cpu-baseline --no-snap) > $out/bench_16M.json 2> $out/bench_16M.err; tail -c 900 $out/bench_16M.json; echo; tail -3 $out/bench_16M.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
