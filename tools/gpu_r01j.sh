#!/bin/bash
# r01j (2 GPUs): TP bench at N=2 as the driver launches it, reference arm, chain test after the fix.
TAG=${1:-r01j}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_smi.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_chain" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "rc=$?"; tail -3 $O/${TAG}_bench_n2.err; cat $O/${TAG}_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --model llama-3-70b > $O/${TAG}_bench_n2_70b.json 2> $O/${TAG}_bench_n2_70b.err; echo "rc=$?"; tail -3 $O/${TAG}_bench_n2_70b.err; cat $O/${TAG}_bench_n2_70b.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --stream 1 --no-cpu-baseline > $O/${TAG}_bench_n2_stream.json 2> $O/${TAG}_bench_n2_stream.err; echo "rc=$?"; tail -3 $O/${TAG}_bench_n2_stream.err; cat $O/${TAG}_bench_n2_stream.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>&1; cat $O/${TAG}_bench_ref.json
for k in skinny; do timeout 300 python tools/microbench.py --quick --kernel $k --ms 8,16 --shapes big 2>&1; done > $O/${TAG}_skinny_big.txt; cat $O/${TAG}_skinny_big.txt
