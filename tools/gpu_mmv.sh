#!/bin/bash
# Focused check of the mmv decode kernel: parity, per-shape sweep, bench, one ncu capture.
TAG=${1:-mmv}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mmv or golden or full_size or auto_dispatch or cuda_graph" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 300 python tools/microbench.py --quick --kernel ${KERN:-mmv8} --ms 1,2,4 > $O/${TAG}_micro.txt 2>&1
cat $O/${TAG}_micro.txt
for v in "GBXQ_MMV8_GRID_MULT=1" "GBXQ_MMV8_GRID_MULT=3" "GBXQ_MMV8_CPW=1"; do
  echo "== $v"; env $v timeout 300 python tools/microbench.py --quick --kernel ${KERN:-mmv8} --ms 1 2>&1 | grep -v shape
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done > $O/${TAG}_variants.txt 2>&1
cat $O/${TAG}_variants.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pdl 0 > $O/${TAG}_bench_nopdl.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pdl 1 > $O/${TAG}_bench_pdl1.json 2>> $O/${TAG}_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmv8 -s 2 -c 1 -f -o $O/${TAG}_k python tools/ncu_one.py 14336 4096 4 64 1 ${KERN:-mmv8} 6 > $O/${TAG}_ncu.log 2>&1
tail -2 $O/${TAG}_ncu.log
timeout 300 python tools/microbench.py --quick --kernel ${KERN:-mmv8} --ms 1 --shapes big > $O/${TAG}_big.txt 2>&1
cat $O/${TAG}_big.txt
