#!/bin/bash
# Recover the measured state: full GPU tests, per-kernel M=1 sweeps, bench with both strategies.
TAG=${1:-r01b}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
for k in gemv mmv mmv8; do
  echo "== $k"; timeout 300 python tools/microbench.py --quick --kernel $k --ms 1,4 2>&1
done > $O/${TAG}_micro_kernels.txt
cat $O/${TAG}_micro_kernels.txt
timeout 300 python tools/microbench.py --quick --kernel mmv8 --ms 1 --shapes big > $O/${TAG}_big.txt 2>&1; cat $O/${TAG}_big.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pdl 0 > $O/${TAG}_bench_nopdl.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_nopdl.json
timeout 600 python bench.py --steps 20 --warmup 5 --strategy bpw-2.2 --no-cpu-baseline > $O/${TAG}_bench_bpw22.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench_bpw22.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmv8 -s 2 -c 1 -f -o $O/${TAG}_mmv8 python tools/ncu_one.py 14336 4096 4 64 1 mmv8 6 > $O/${TAG}_ncu.log 2>&1
tail -2 $O/${TAG}_ncu.log
