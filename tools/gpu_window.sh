# the driver's window (--steps 20 --warmup 5) against the steady state, at N GPUs
n=${1:-2}; tag=${2:-r02win}
out=gpurun_out/$tag; mkdir -p $out
for a in "20 5" "200 40"; do
  set -- $a
  if [ $n -gt 1 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps $1 --warmup $2 --no-extra --no-cpu-baseline --no-parity > $out/bench_n${n}_$1.json 2> $out/bench_n${n}_$1.err
  else
    timeout 600 python bench.py --steps $1 --warmup $2 --no-extra --no-cpu-baseline > $out/bench_n${n}_$1.json 2> $out/bench_n${n}_$1.err
  fi
  python - <<PY
import json
d=json.loads(open('$out/bench_n${n}_$1.json').read().strip().splitlines()[-1])
print('n=$n steps $1 warmup $2', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'no_thermo %.4f'%d.get('ms_per_step_no_thermo',0), {k: round(v,4) for k,v in d['phase_ms_per_step'].items()})
PY
done
