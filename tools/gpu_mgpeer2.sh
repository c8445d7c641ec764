n=${1:-2}; tag=${2:-r02t}; shift; shift
out=gpurun_out/$tag; mkdir -p $out
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 200 --warmup 40 --no-extra --no-cpu-baseline --no-parity > $out/bench_n${n}_$name.json 2> $out/bench_n${n}_$name.err
  grep emd_peer $out/bench_n${n}_$name.err | head -2
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n${n}_$name.json').read().strip().splitlines()[-1])
    print('n=$n $name', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'no_thermo %.4f'%d.get('ms_per_step_no_thermo',0), {k: round(v,4) for k,v in d['phase_ms_per_step'].items()})
except Exception as e:
    print('n=$n $name FAILED', e); print(open('$out/bench_n${n}_$name.err').read()[-2500:])
PY
}
run peer_noov EMD_HALO_TRANSPORT=peer EMD_NO_OVERLAP=1 EMD_PEER_DEBUG=1
run peer_ov EMD_HALO_TRANSPORT=peer
run nccl_ov EMD_HALO_TRANSPORT=nccl
