#!/bin/bash
TAG=${1:-r01h}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_chain" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 200 python tools/stream_timeline.py 8 > $O/${TAG}_stream_timeline.txt 2>&1; head -14 $O/${TAG}_stream_timeline.txt; tail -7 $O/${TAG}_stream_timeline.txt
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
{
for v in "--stream 1" "--stream 1 --strategy bpw-2.2" "--stream 1 --batch 2"; do
  echo "== $v"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v 2>&1 | tail -3 | b
done
} > $O/${TAG}_bench.txt 2>&1
cat $O/${TAG}_bench.txt | tail -30
