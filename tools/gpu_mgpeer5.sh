n=${1:-2}; tag=${2:-r02x}
out=gpurun_out/$tag; mkdir -p $out
EMD_PEER_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 200 --warmup 40 --no-extra --no-cpu-baseline --no-parity > $out/bench.json 2> $out/bench.err
grep "emd_peer\|CommMPI\[" $out/bench.err; tail -c 300 $out/bench.json
