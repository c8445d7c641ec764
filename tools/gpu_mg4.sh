# 4-GPU check (two decomposed dimensions): one parity case + the weak-scaling bench line; tight timeouts
out=gpurun_out/${1:-r01y}; mkdir -p $out
MASTER_ADDR=127.0.0.1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py lj 12 14 14 45 half 2>&1 | grep "^MGPU" | tail -3
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --steps 100 --warmup 20 > $out/bench_n4.json 2> $out/bench_n4.err
python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n4.json').read().strip().splitlines()[-1])
    print('n=4', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, d['config']['parallelism'])
except Exception as e:
    print('FAILED', e); print(open('$out/bench_n4.err').read()[-1500:])
PY
