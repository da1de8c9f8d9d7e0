#!/bin/bash
# One gpurun call for kernel work on the decode path: mmv8 parity tests, per-shape microbench, default bench, timelines.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_kernel.sh r02a'
TAG=${1:-k}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 200 python bench.py --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'])"
for shp in "14336 4096 4 64" "4096 14336 4 64" "4096 4096 4 64"; do echo "== timeline $shp"; timeout 100 python tools/timeline.py $shp 8 2; done > $O/${TAG}_timeline.txt 2>&1
timeout 120 python tools/microbench.py --quick --kernel mmv8 --ms 1 > $O/${TAG}_micro.txt 2>&1
timeout 120 python tools/microbench.py --quick --kernel mmv8 --ms 1 --shapes big >> $O/${TAG}_micro.txt 2>&1
timeout 120 python tools/microbench.py --quick --kernel mmv8 --ms 1 --l2 >> $O/${TAG}_micro.txt 2>&1
cat $O/${TAG}_micro.txt
