# single-GPU check: parity suite, bench line (ELL ring on/off), launch list
tag=${1:-r01t}
mkdir -p gpurun_out/$tag
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for ring in 0 1; do
EMD_TILES_NO_RING=$ring timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-snap > gpurun_out/$tag/bench_noring$ring.json 2> gpurun_out/$tag/bench_noring$ring.err
python - <<PY
import json
d=json.load(open('gpurun_out/$tag/bench_noring$ring.json'))
print('no_ring=$ring value %.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'force_ms %.4f'%d['roofline']['kernel_ms'], d['phase_ms_per_step'], 'e2e %.3e'%d['e2e']['value'])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lj_tiles_kernel" -s 2 -c 1 -o gpurun_out/$tag/lj_ring -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snap > gpurun_out/$tag/ncu.log 2>&1
