#!/bin/bash
# r01g: chain launch after the descriptor-prefetch and cost-balance fixes; M = 2..4 kernel choice; timeline.
TAG=${1:-r01g}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_chain or grouped" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 200 python tools/stream_timeline.py 8 > $O/${TAG}_stream_timeline.txt 2>&1; tail -8 $O/${TAG}_stream_timeline.txt
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
{
for v in "--stream 1" "--stream 0" "--stream 1 --strategy bpw-2.2" "--stream 0 --strategy bpw-2.2" "--stream 1 --batch 2" "--stream 0 --batch 2" "--stream 0 --batch 4" "--stream 0 --batch 8" "--stream 0 --batch 16"; do
  echo "== $v"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v 2>&1 | tail -1 | b
done
} > $O/${TAG}_bench.txt 2>&1
grep -v "^  File\|^    \|Traceback\|json" $O/${TAG}_bench.txt
for k in mmv8 skinny; do echo "== $k"; timeout 300 python tools/microbench.py --quick --kernel $k --ms 2,3,4 2>&1 | grep -v "^shape" ; done > $O/${TAG}_micro_m234.txt 2>&1
cat $O/${TAG}_micro_m234.txt
