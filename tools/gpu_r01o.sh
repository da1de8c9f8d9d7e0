#!/bin/bash
# r01o: end-to-end decode through the callers (checkpoint -> load -> generate), configs[0] and 3B.
TAG=${1:-r01o}
O=gpurun_out
mkdir -p $O
timeout 300 python tools/decode_bench.py --model llama-3.2-1b > $O/${TAG}_decode_1b.json 2> $O/${TAG}_decode_1b.err; echo "rc=$?"; tail -2 $O/${TAG}_decode_1b.err; cat $O/${TAG}_decode_1b.json
timeout 400 python tools/decode_bench.py --model llama-3.2-3b > $O/${TAG}_decode_3b.json 2> $O/${TAG}_decode_3b.err; echo "rc=$?"; tail -2 $O/${TAG}_decode_3b.err; cat $O/${TAG}_decode_3b.json
