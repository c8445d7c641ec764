out=gpurun_out/r02win4; mkdir -p $out
for kw in "20 5" "100 20"; do set -- $kw
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps $1 --warmup $2 --no-extra --no-cpu-baseline > $out/bench_n4_$1_$2.json 2> $out/bench_n4_$1_$2.err
python - <<PY
import json
d=json.loads(open('$out/bench_n4_$1_$2.json').read().strip().splitlines()[-1])
print('K=$1 W=$2', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'no_thermo %.4f'%d.get('ms_per_step_no_thermo',0), {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, d.get('parity',{}).get('ok'))
PY
done
