# single-GPU check: parity suite, bench line with and without the force+integrator fusion
tag=${1:-r01w}
mkdir -p gpurun_out/$tag
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for off in 0 1; do
EMD_NO_FUSED_FORCE_NVE=$off timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-snap > gpurun_out/$tag/bench_nofuse$off.json 2> gpurun_out/$tag/bench_nofuse$off.err
python - <<PY
import json
d=json.load(open('gpurun_out/$tag/bench_nofuse$off.json'))
print('no_force_fuse=$off value %.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'force_ms %.4f'%d['roofline']['kernel_ms'], d['phase_ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'launches', d['gpu_launches'])
PY
done
