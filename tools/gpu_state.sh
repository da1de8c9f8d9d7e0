#!/bin/bash
# Re-measure the state of the repo in ONE gpurun call (~6 GPU-minutes): tests, smoke, default bench, variants, the two
# timelines, asymptotic rates, end-to-end decode.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_state.sh r02a'
TAG=${1:-state}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
for v in "--stream 1" "--strategy bpw-2.2" "--batch 2" "--batch 4" "--batch 16" "--model llama-3-70b --steps 5"; do
  echo "== $v"; timeout 200 python bench.py --no-cpu-baseline $v 2>&1 | tail -1 | b
done > $O/${TAG}_bench_variants.txt 2>&1; cat $O/${TAG}_bench_variants.txt
for shp in "14336 4096 4 64" "4096 14336 4 64"; do echo "== timeline $shp"; timeout 100 python tools/timeline.py $shp 8 2; done > $O/${TAG}_timeline.txt 2>&1
timeout 150 python tools/stream_timeline.py 8 > $O/${TAG}_stream_timeline.txt 2>&1; tail -7 $O/${TAG}_stream_timeline.txt
for k in mmv8 skinny; do echo "== $k big"; timeout 120 python tools/microbench.py --quick --kernel $k --ms $([ $k = mmv8 ] && echo 1 || echo 8) --shapes big 2>&1 | grep -v "^shape"; done > $O/${TAG}_big.txt 2>&1; cat $O/${TAG}_big.txt
timeout 150 python tools/decode_bench.py --model llama-3.2-1b --repeat 2 > $O/${TAG}_decode_1b.json 2> $O/${TAG}_decode_1b.err; cat $O/${TAG}_decode_1b.json
