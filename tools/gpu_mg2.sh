n=${1:-2}; tag=${2:-mg}
out=gpurun_out/$tag; mkdir -p $out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 100 --warmup 20 > $out/bench_n${n}_$name.json 2> $out/bench_n${n}_$name.err
  python - <<PY
import json
d=json.loads(open('$out/bench_n${n}_$name.json').read().strip().splitlines()[-1])
print('n=$n $name', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phase_ms_per_step'].items()})
PY
}
run noov EMD_NO_OVERLAP=1
run forcefirst_r16 EMD_OVERLAP_ORDER=0 EMD_OVERLAP_RESERVE=16
run commfirst_r16 EMD_OVERLAP_ORDER=1 EMD_OVERLAP_RESERVE=16
run commfirst_r0 EMD_OVERLAP_ORDER=1 EMD_OVERLAP_RESERVE=0
run commfirst_r148 EMD_OVERLAP_ORDER=1 EMD_OVERLAP_RESERVE=148
run forcefirst_r148 EMD_OVERLAP_ORDER=0 EMD_OVERLAP_RESERVE=148
