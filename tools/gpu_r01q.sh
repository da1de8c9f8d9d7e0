#!/bin/bash
# r01q: final state of the round: all GPU tests, smoke, default bench, end-to-end decode with the PDL-aware glue.
TAG=${1:-r01t}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -2 $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
for m in llama-3.2-1b llama-3-8b; do
  timeout 300 python tools/decode_bench.py --model $m > $O/${TAG}_decode_$m.json 2> $O/${TAG}_decode_$m.err; echo "$m rc=$?"; tail -2 $O/${TAG}_decode_$m.err; cat $O/${TAG}_decode_$m.json
done
