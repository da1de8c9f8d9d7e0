#!/bin/bash
# r01p: decode-step glue kernels: unit + model tests, end-to-end decode with and without them.
TAG=${1:-r01p}
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_glue.py tests/test_gpu_model.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -25 $O/${TAG}_pytest.log
for m in llama-3.2-1b llama-3.2-3b llama-3-8b; do
  timeout 300 python tools/decode_bench.py --model $m > $O/${TAG}_decode_$m.json 2> $O/${TAG}_decode_$m.err; echo "$m rc=$?"; tail -2 $O/${TAG}_decode_$m.err; cat $O/${TAG}_decode_$m.json
done
GBXQ_FUSED_DECODE=0 timeout 300 python tools/decode_bench.py --model llama-3-8b > $O/${TAG}_decode_8b_unfused.json 2> $O/${TAG}_decode_8b_unfused.err; cat $O/${TAG}_decode_8b_unfused.json
