#!/bin/bash
N=${1:-2}
TAG=r01n
O=gpurun_out
mkdir -p $O
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err; echo "rc=$?"; tail -4 $O/${TAG}_n$N.err | cut -c1-400; cat $O/${TAG}_n$N.json
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus $N --steps 5 --warmup 3 --model llama-3-70b --no-also > $O/${TAG}_n${N}_70b.json 2> $O/${TAG}_n${N}_70b.err; echo "rc=$?"; tail -2 $O/${TAG}_n${N}_70b.err | cut -c1-400; cat $O/${TAG}_n${N}_70b.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err; echo "rc=$?"; tail -2 $O/${TAG}_n1.err; cut -c1-400 $O/${TAG}_n1.json
