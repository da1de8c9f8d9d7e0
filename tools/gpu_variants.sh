#!/bin/bash
# bench.py under environment / flag variants, one line each.  Usage: gpurun -- 'bash tools/gpu_variants.sh TAG'
TAG=${1:-v}
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
run() { echo "== $*"; env $1 timeout 200 python bench.py --no-cpu-baseline $2 2>&1 | tail -1 | b; }
{
run "X=1" ""
run "GBXQ_MMV8_STAGES=6" ""
run "GBXQ_MMV8_STAGES=8 GBXQ_MMV8_RING_KB=100" ""
run "GBXQ_MMV8_STAGES=3" ""
run "X=1" "--stream 1"
run "X=1" "--strategy bpw-2.2"
run "X=1" "--batch 2"
run "X=1" "--batch 4"
run "X=1" "--model llama-3-70b --steps 5"
run "X=1" "--model llama-3.2-3b"
} > $O/${TAG}_variants.txt 2>&1
cat $O/${TAG}_variants.txt
