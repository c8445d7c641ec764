# single-GPU check: parity suite, bench line, launch list
tag=${1:-r01s}
mkdir -p gpurun_out/$tag
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-snap > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/$tag/bench.json'))
print('value %.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'force_ms %.4f'%d['roofline']['kernel_ms'], d['phase_ms_per_step'], 'e2e %.3e'%d['e2e']['value'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$tag/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snap > gpurun_out/$tag/ncu_bench.log 2>&1
