#!/bin/bash
TAG=${1:-r01s}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -12 $O/${TAG}_pytest.log
for m in llama-3.2-1b llama-3-8b; do
  timeout 240 python tools/decode_bench.py --model $m --repeat 2 > $O/${TAG}_decode_$m.json 2> $O/${TAG}_decode_$m.err; echo "$m rc=$?"; tail -3 $O/${TAG}_decode_$m.err; cat $O/${TAG}_decode_$m.json
done
