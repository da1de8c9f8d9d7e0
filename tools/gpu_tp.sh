#!/bin/bash
# Multi-GPU run: TP tests, then bench.py at N GPUs with each all-reduce flavour.  Usage: gpurun --gpus N -- 'bash tools/gpu_tp.sh TAG N [model]'
TAG=${1:-tp}; N=${2:-2}; MODEL=${3:-llama-3-8b}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tp.py -m gpu -x -q -s > $O/${TAG}_pytest_tp.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_tp.log; tail -5 $O/${TAG}_pytest_tp.log
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --model $MODEL "$@" 2> $O/${TAG}_bench_n${N}.err | tail -1; }
for ar in fused oneshot nccl; do
  echo "== allreduce $ar"; run --allreduce $ar $([ $ar != fused ] && echo --no-also) | tee $O/${TAG}_bench_n${N}_${ar}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['parallelism'], d['config']['allreduce'][:60], '| also', d.get('also'))"
done
