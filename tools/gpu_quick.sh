#!/bin/bash
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 300 python tools/microbench.py --quick --kernel mmv8 --ms 1 > $O/${TAG}_micro.txt 2>&1
cat $O/${TAG}_micro.txt
timeout 300 python tools/microbench.py --quick --kernel mmv8 --ms 1 --shapes big 2>&1 | grep -v shape
for v in "X=1" "GBXQ_MMV8_GRID_MULT=1" "GBXQ_MMV8_GRID_MULT=3" "GBXQ_MMV8_RING_KB=40"; do
  echo "== $v"; env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pdl 0 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench nopdl', d['value'], d['ms_per_step'], d['roofline']['frac'])"
