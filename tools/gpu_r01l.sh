#!/bin/bash
# r01l (2 GPUs): TP bench as the driver launches it: one-shot all-reduce vs NCCL, clean exit.
TAG=${1:-r01l}
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 "$@" > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err; echo "$name rc=$?"; tail -2 $O/${TAG}_$name.err | cut -c1-300; cut -c1-700 $O/${TAG}_$name.json; }
run n2 --steps 20 --warmup 5
run n2_nccl --steps 20 --warmup 5 --allreduce nccl
run n2_70b --steps 5 --warmup 3 --model llama-3-70b
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>&1; cut -c1-600 $O/${TAG}_bench_ref.json
