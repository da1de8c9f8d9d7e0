#!/bin/bash
# r01i: full state after the chain launch / cost balance / prologue / dispatch changes.
TAG=${1:-r01i}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
{
for v in "--stream 1" "--strategy bpw-2.2" "--batch 2" "--batch 4" "--batch 8" "--batch 16" "--model llama-3.2-3b" "--model llama-3-70b --steps 5"; do
  echo "== $v"; timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v 2>&1 | tail -1 | b
done
} > $O/${TAG}_bench_variants.txt 2>&1
grep -v "^  File\|^    \|Traceback\|json" $O/${TAG}_bench_variants.txt
for shp in "4096 14336 4 64" "14336 4096 4 64"; do echo "== timeline $shp"; timeout 120 python tools/timeline.py $shp 8 2; done > $O/${TAG}_timeline.txt 2>&1
cat $O/${TAG}_timeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmv8_grouped -s 41 -c 1 -f -o $O/${TAG}_mmv8g python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_mmv8g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmv8_kernel -s 2 -c 1 -f -o $O/${TAG}_mmv8 python tools/ncu_one.py 14336 4096 4 64 1 mmv8 6 > $O/${TAG}_ncu_mmv8.log 2>&1
ls -la $O | tail -20
