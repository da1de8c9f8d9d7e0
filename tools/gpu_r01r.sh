#!/bin/bash
TAG=${1:-r01r}
O=gpurun_out
mkdir -p $O
for m in llama-3.2-1b llama-3-8b; do
  timeout 240 python tools/decode_bench.py --model $m --repeat 2 > $O/${TAG}_decode_$m.json 2> $O/${TAG}_decode_$m.err; echo "$m rc=$?"; tail -2 $O/${TAG}_decode_$m.err; cat $O/${TAG}_decode_$m.json
done
