#!/bin/bash
# decode-kernel geometry sweep: bench.py per (environment, workload)
TAG=${1:-sweep}; O=gpurun_out; mkdir -p $O
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'])"; }
{
for e in "GBXQ_MMV8_TWOCOLS=16" "GBXQ_MMV8_TWOCOLS=16 GBXQ_MMV8_R16=0" "GBXQ_MMV8_TWOCOLS=8" "GBXQ_MMV8_TWOCOLS=32 GBXQ_MMV8_R16=0"; do
  for v in "" "--model llama-3.2-3b" "--strategy bpw-2.2" "--model llama-3-70b --steps 5" "--batch 2"; do
    echo "== [$e] $v"; env $e timeout 200 python bench.py --no-cpu-baseline $v 2>&1 | tail -1 | b
  done
done
} > $O/${TAG}_sweep.txt 2>&1
cat $O/${TAG}_sweep.txt
