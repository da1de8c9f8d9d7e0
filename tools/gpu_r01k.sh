#!/bin/bash
# r01k (1 GPU): full GPU tests, default bench, ncu launch list of OUR kernels over one step, DRAM traffic per launch.
TAG=${1:-r01k}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
for v in "--batch 4" "--batch 8"; do echo "== $v"; timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v 2>&1 | tail -1 | b; done > $O/${TAG}_bench_variants.txt 2>&1; cat $O/${TAG}_bench_variants.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mmv8 -c 384 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
python tools/traffic_from_ncu.py $O/${TAG}_launches.csv 128 $O/${TAG}_traffic.json
timeout 120 python tools/microbench.py --quick --kernel skinny --ms 8 --shapes big > $O/${TAG}_skinny_big.txt 2>&1; cat $O/${TAG}_skinny_big.txt
