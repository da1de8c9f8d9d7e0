#!/bin/bash
# r01f: chain launch (gbxq_qmm_stream): parity tests, bench against the launch-per-call step, stage-count sweep.
TAG=${1:-r01f}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_chain or grouped" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'], d['config'].get('chain_launch'))"; }
{
for v in "--stream 1" "--stream 0" "--stream 1 --grouped 0" "--stream 1 --strategy bpw-2.2" "--stream 1 --batch 2" "--stream 1 --batch 4"; do
  echo "== $v"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v 2>&1 | tail -1 | b
done
for st in 2 3 4; do
  echo "== stages $st"; GBXQ_STREAM_STAGES=$st timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | b
done
} > $O/${TAG}_bench.txt 2>&1
cat $O/${TAG}_bench.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -2 $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
