# multi-GPU experiment matrix: bash tools/gpu_mg2.sh <ngpus> <tag> ; each line = name + environment
n=${1:-2}; tag=${2:-mg}
out=gpurun_out/$tag; mkdir -p $out
run() {
  name=$1; shift
  env "$@" EMD_VERBOSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 100 --warmup 20 > $out/bench_n${n}_$name.json 2> $out/bench_n${n}_$name.err
  grep -m1 "side stream" $out/bench_n${n}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/bench_n${n}_$name.json').read().strip().splitlines()[-1])
    print('n=$n $name', 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phase_ms_per_step'].items()})
except Exception as e:
    print('n=$n $name FAILED', e); print(open('$out/bench_n${n}_$name.err').read()[-1500:])
PY
}
run noov EMD_NO_OVERLAP=1
run noov_ce EMD_NO_OVERLAP=1 NCCL_P2P_USE_CUDA_MEMCPY=1
run part12_ce EMD_OVERLAP_SMS=12 NCCL_P2P_USE_CUDA_MEMCPY=1
run plain_ce EMD_OVERLAP_SMS=0 NCCL_P2P_USE_CUDA_MEMCPY=1
run noov_ch32 EMD_NO_OVERLAP=1 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
