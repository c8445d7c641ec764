# round-2 final parity evidence on a 2-GPU box: the multi-GPU tests (CommMPI cases, all halo schedules), then the 2-GPU bench line
out=gpurun_out/r02c; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s -rs > $out/pytest_multi.log 2>&1; echo "pytest rc $?" >> $out/pytest_multi.log
tail -5 $out/pytest_multi.log; grep MGPU $out/pytest_multi.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 100 --warmup 20 > $out/bench_n2.json 2> $out/bench_n2.err
tail -c 2500 $out/bench_n2.json
