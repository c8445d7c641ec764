n=${1:-2}; tag=${2:-r02w}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s -rs > $out/pytest_multi.log 2>&1; tail -3 $out/pytest_multi.log
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 200 --warmup 40 --no-extra --no-cpu-baseline > $out/bench_n${n}_$name.json 2> $out/bench_n${n}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n${n}_$name.json').read().strip().splitlines()[-1])
    print('n=$n $name', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'no_thermo %.4f'%d.get('ms_per_step_no_thermo',0), {k: round(v,4) for k,v in d['phase_ms_per_step'].items()}, d.get('parity',{}).get('ok'))
except Exception as e:
    print('n=$n $name FAILED', e); print(open('$out/bench_n${n}_$name.err').read()[-2500:])
PY
}
run gated EMD_HALO_TRANSPORT=peer EMD_HALO_GATE=1
run peer_split EMD_HALO_TRANSPORT=peer
