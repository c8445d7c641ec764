# round-2 parity evidence: full -m gpu suite on a 2-GPU box (CommMPI cases run), then the 2-GPU bench line with its parity block
out=gpurun_out/r02b; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x -s -rs > $out/pytest.log 2>&1; echo "pytest rc $?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 100 --warmup 20 > $out/bench_n2.json 2> $out/bench_n2.err
tail -c 1500 $out/bench_n2.json
