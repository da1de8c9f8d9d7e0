#!/bin/bash
# r01m (N GPUs): the driver's launch at N = 4 / 8, tight limits.
N=${1:-8}
TAG=r01m
O=gpurun_out
mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err; echo "rc=$?"; tail -4 $O/${TAG}_n$N.err | cut -c1-400; cut -c1-900 $O/${TAG}_n$N.json
