# final 8-GPU bench line of round 2 (headline 2 M atoms per GPU + parity block + weak16M + snap_strong sections)
out=gpurun_out/r02n8; mkdir -p $out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 100 --warmup 20 --no-cpu-baseline > $out/bench_n8.json 2> $out/bench_n8.err
python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n8.json').read().strip().splitlines()[-1])
    print('n=8', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, 'parity', d.get('parity',{}).get('ok'))
    for k in ('weak16M','snap_strong'):
        if k in d: print(k, 'value %.4e'%d[k]['value'], 'ms/step %.3f'%d[k]['ms_per_step'], {a: round(b,3) for a,b in d[k]['phase_ms_per_step'].items()}, d[k].get('roofline',{}).get('frac'))
except Exception as e:
    print('FAILED', e); print(open('$out/bench_n8.err').read()[-2500:])
PY
