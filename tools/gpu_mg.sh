# multi-GPU check: parity tests + weak-scaling bench with and without halo/force overlap
n=${1:-2}; tag=${2:-mg}
out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiles" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for ov in 1 0; do
EMD_NO_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$ov bench.py --gpus $n --steps 100 --warmup 20 > $out/bench_n${n}_noov$ov.json 2> $out/bench_n${n}_noov$ov.err
python - <<PY
import json
d=json.loads(open('$out/bench_n${n}_noov$ov.json').read().strip().splitlines()[-1])
print('n=$n no_overlap=$ov', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['phase_ms_per_step'])
PY
done
