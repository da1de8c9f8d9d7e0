#!/bin/bash
# default bench + per-shape sweep of one library build under several environment settings.
# Usage: gpurun -- 'bash tools/gpu_envs.sh TAG lib.so "ENV=1 ENV2=2" "ENV=3" ...'
TAG=${1:-envs}; LIB=$2; shift; shift
O=gpurun_out; mkdir -p $O
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
export GBXQ_LIB=$PWD/gbx-lm_b200/$LIB
{
for envs in "$@"; do
  echo "#### $LIB env[$envs]"
  env $envs timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | b
  env $envs timeout 200 python tools/microbench.py --quick --kernel mmv8 --ms 1 2>&1 | grep -v "^shape"
done
} > $O/${TAG}_envs.txt 2>&1
cat $O/${TAG}_envs.txt
