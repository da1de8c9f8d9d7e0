#!/bin/bash
# State of the repo in one call: GPU tests, smoke, default bench, decode-batch / prefill / model variants, end-to-end decode.
TAG=${1:-state}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -6 $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['unit'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
for v in "--strategy bpw-2.2" "--batch 2" "--batch 4" "--batch 8" "--batch 16" "--batch 32" "--batch 64" "--model llama-3.2-3b" "--model llama-3-70b --steps 5" "--phase prefill" "--phase prefill --model llama-3.2-3b" "--model qwen2.5-32b --strategy uniform --group-size 128 --batch 32 --steps 5"; do
  echo "== $v"; timeout 300 python bench.py --no-cpu-baseline $v 2>&1 | tail -1 | b
done > $O/${TAG}_bench_variants.txt 2>&1; cat $O/${TAG}_bench_variants.txt
timeout 200 python tools/decode_bench.py --model llama-3-8b --repeat 2 > $O/${TAG}_decode_8b.json 2> $O/${TAG}_decode_8b.err; cat $O/${TAG}_decode_8b.json
