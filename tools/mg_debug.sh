for cfg in "lj 12 10 12 45 half" "snap 4 4 8 6"; do
MASTER_ADDR=127.0.0.1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py $cfg 2>&1 | grep "^MGPU"
done
