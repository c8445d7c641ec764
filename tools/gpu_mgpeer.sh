# peer-store halo refresh vs NCCL send/recv groups: parity tests + bench lines   bash tools/gpu_mgpeer.sh <ngpus> <tag>
n=${1:-2}; tag=${2:-r02s}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s -rs > $out/pytest_multi.log 2>&1; tail -4 $out/pytest_multi.log
for tr in peer nccl; do
  EMD_HALO_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 200 --warmup 40 --no-extra --no-cpu-baseline > $out/bench_n${n}_$tr.json 2> $out/bench_n${n}_$tr.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n${n}_$tr.json').read().strip().splitlines()[-1])
    print('n=$n $tr', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, d.get('parity',{}).get('ok'))
except Exception as e:
    print('n=$n $tr FAILED', e); print(open('$out/bench_n${n}_$tr.err').read()[-2500:])
PY
done
