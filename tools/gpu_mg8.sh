# 8-GPU weak-scaling bench line (three decomposed dimensions); tight timeout
out=gpurun_out/${1:-r01z8}; mkdir -p $out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 100 --warmup 20 > $out/bench_n8.json 2> $out/bench_n8.err
python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n8.json').read().strip().splitlines()[-1])
    print('n=8', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, d['config']['parallelism'])
except Exception as e:
    print('FAILED', e); print(open('$out/bench_n8.err').read()[-1500:])
PY
