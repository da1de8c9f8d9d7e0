#!/bin/bash
# r01c: launch-chain floor + L2 prefetch probes, CTA co-residency variants of the decode kernel, timelines.
TAG=${1:-r01c}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 300 tools/ubench/ub_chain > $O/${TAG}_chain.txt 2>&1; cat $O/${TAG}_chain.txt
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'])"; }
{
for v in "X=1" "GBXQ_MMV8_GRID_MULT=1" "GBXQ_MMV8_GRID_MULT=1 GBXQ_MMV8_RING_KB=64" "GBXQ_MMV8_GRID_MULT=1 GBXQ_MMV8_RING_KB=48 GBXQ_MMV8_STAGE_KB=16" "GBXQ_MMV8_RING_KB=48 GBXQ_MMV8_STAGE_KB=16" "GBXQ_MMV8_GRID_MULT=3 GBXQ_MMV8_RING_KB=64 GBXQ_MMV8_STAGE_KB=16"; do
  echo "== $v"; env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | b
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --strategy bpw-2.2 | b
done
} > $O/${TAG}_variants.txt 2>&1
cat $O/${TAG}_variants.txt
{
for v in "X=1" "GBXQ_MMV8_GRID_MULT=1"; do
  echo "== $v big"; env $v timeout 300 python tools/microbench.py --quick --kernel mmv8 --ms 1 --shapes big 2>&1 | grep -v shape
  echo "== $v 8b"; env $v timeout 300 python tools/microbench.py --quick --kernel mmv8 --ms 1 2>&1 | grep -v shape
  for shp in "14336 4096 4 64" "1024 4096 4 64" "4096 4096 4 64"; do
    echo "== $v timeline $shp"; env $v timeout 120 python tools/timeline.py $shp 8 2
  done
done
} > $O/${TAG}_timeline.txt 2>&1
cat $O/${TAG}_timeline.txt
