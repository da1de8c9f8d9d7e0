#!/bin/bash
# bench.py under environment / flag variants, one line each.  Usage: gpurun -- 'bash tools/gpu_variants.sh TAG "ENV1|FLAGS1" "ENV2|FLAGS2" ...'
TAG=${1:-v}; shift
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
for spec in "$@"; do
  envs="${spec%%|*}"; flags="${spec#*|}"
  echo "== env[$envs] flags[$flags]"
  env $envs timeout 200 python bench.py --no-cpu-baseline $flags 2>&1 | tail -1 | b
done > $O/${TAG}_variants.txt 2>&1
cat $O/${TAG}_variants.txt
