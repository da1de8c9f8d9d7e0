#!/bin/bash
TAG=${1:-r01u}
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_glue.py tests/test_gpu_model.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -8 $O/${TAG}_pytest.log
timeout 100 python tools/decode_bench.py --model llama-3.2-1b --repeat 2 > $O/${TAG}_decode_1b.json 2> $O/${TAG}_decode_1b.err; echo "rc=$?"; tail -2 $O/${TAG}_decode_1b.err; cat $O/${TAG}_decode_1b.json
