n=${1:-2}; tag=${2:-r02pair}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s -rs > $out/pytest_multi.log 2>&1; tail -3 $out/pytest_multi.log
EMD_PEER_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 200 --warmup 40 --no-extra --no-cpu-baseline > $out/bench_n$n.json 2> $out/bench_n$n.err
grep "CommMPI\[" $out/bench_n$n.err
python - <<PY
import json
d=json.loads(open('$out/bench_n$n.json').read().strip().splitlines()[-1])
print('n=$n', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'no_thermo %.4f'%d.get('ms_per_step_no_thermo',0), {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, d.get('parity',{}).get('ok'))
PY
