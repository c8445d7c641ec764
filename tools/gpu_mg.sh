# multi-GPU check: parity tests + weak-scaling bench line (short timeouts: a hang must not burn the GPU budget)
n=${1:-2}; tag=${2:-mg}
out=gpurun_out/$tag; mkdir -p $out
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for off in 0 1; do
EMD_NO_FUSED_FORCE_NVE=$off timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$off bench.py --gpus $n --steps 100 --warmup 20 > $out/bench_n${n}_nofuse$off.json 2> $out/bench_n${n}_nofuse$off.err
python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n${n}_nofuse$off.json').read().strip().splitlines()[-1])
    print('n=$n no_force_fuse=$off', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, 'lines', len(open('$out/bench_n${n}_nofuse$off.json').read().strip().splitlines()))
except Exception as e:
    print('FAILED', e); print(open('$out/bench_n${n}_nofuse$off.err').read()[-1200:])
PY
done
