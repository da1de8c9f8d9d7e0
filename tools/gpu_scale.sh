#!/bin/bash
# TP scaling measurements on one multi-GPU box.  Usage: gpurun --gpus G -- 'bash tools/gpu_scale.sh TAG "model:N:extra flags" ...'
TAG=${1:-scale}; shift
O=gpurun_out
mkdir -p $O
for spec in "$@"; do
  model="${spec%%:*}"; rest="${spec#*:}"; n="${rest%%:*}"; flags="${rest#*:}"; [ "$flags" = "$rest" ] && flags=""
  name=$(echo "${model}_n${n}_${flags}" | tr -c 'A-Za-z0-9_.\n-' '_')
  echo "== $model N=$n $flags"
  if [ "$n" = "1" ]; then
    timeout 240 python bench.py --model $model --no-cpu-baseline --steps 5 --warmup 3 $flags 2> $O/${TAG}_${name}.err | tail -1 > $O/${TAG}_${name}.json
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus $n --model $model --steps 5 --warmup 3 --no-also $flags 2> $O/${TAG}_${name}.err | tail -1 > $O/${TAG}_${name}.json
  fi
  python -c "import json,sys; d=json.load(open('$O/${TAG}_${name}.json')); print(d['value'], d['unit'], d['ms_per_step'], 'ms', d['config']['parallelism'], '| tok/s', d.get('decode_tok_s_qmm_only'), '| frac', d['roofline']['frac'], '|', d['config']['allreduce'][:50])" || tail -3 $O/${TAG}_${name}.err
done 2>&1 | tee $O/${TAG}_summary.txt
