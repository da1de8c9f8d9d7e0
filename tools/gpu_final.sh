#!/bin/bash
# End-of-round validation in one call: the GPU suite, smoke, the default bench line (with its 'also' legs), the reference
# arm, the end-to-end decode.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_final.sh TAG'
TAG=${1:-final}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; cat $O/${TAG}_bench_reference.json
timeout 200 python tools/decode_bench.py --model llama-3-8b --repeat 2 > $O/${TAG}_decode_8b.json 2> $O/${TAG}_decode_8b.err; cat $O/${TAG}_decode_8b.json
