# round-2 evidence on one B200: default bench line, driver-style short bench line, reference arm, launch list, full ncu captures
tag=${1:-r02fin}
out=gpurun_out/$tag; mkdir -p $out
(time timeout 900 python bench.py) > $out/bench.json 2> $out/bench.err; tail -c 400 $out/bench.json; echo
(time timeout 600 python bench.py --steps 20 --warmup 5) > $out/bench_20_5.json 2> $out/bench_20_5.err; tail -c 300 $out/bench_20_5.json; echo
(time timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 300 $out/bench_reference.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_bench.log 2>&1
# full captures: one re-neighboring (tile kernels, binning, halo) + the force kernels of the following steps
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tiles_|bin_|permute|halo_|lj_tiles' -s 40 -c 40 -o $out/step_full -f \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_full.log 2>&1
ncu -i $out/step_full.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
python tools/ncu_summary.py $out/raw.csv > $out/summary.csv
rm -f $out/step_full.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls -la $out
