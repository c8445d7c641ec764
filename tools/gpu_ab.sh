mkdir -p gpurun_out/r01k
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in rot; do
EMD_TILES_SCHED=$v timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-snap > gpurun_out/r01k/bench_$v.json 2> gpurun_out/r01k/bench_$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/r01k/bench_$v.json'))
print('sched=$v', 'value %.3e'%d['value'], 'force_ms %.4f'%d['roofline']['kernel_ms'], d['phase_ms_per_step'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01k/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snap > gpurun_out/r01k/ncu_bench.log 2>&1
