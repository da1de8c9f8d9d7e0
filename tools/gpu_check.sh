#!/bin/bash
# One gpurun call: GPU tests, bench, roofline sweeps, ncu launch list and one full capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
python -c "import mlx.core" > $O/${TAG}_mlx_probe.txt 2>&1; echo "mlx import rc=$?" >> $O/${TAG}_mlx_probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --strategy bpw-2.2 --no-cpu-baseline > $O/${TAG}_bench_bpw22.json 2>> $O/${TAG}_bench.err
timeout 600 python tools/microbench.py --ms 1,2,4,8,16 --json $O/${TAG}_micro.json > $O/${TAG}_micro.txt 2>&1
timeout 600 python tools/gemmbench.py --ms 32,64,128,512,2048 --bits 4,2 --json $O/${TAG}_gemm.json > $O/${TAG}_gemm.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv -s 2 -c 2 -f -o $O/${TAG}_gemv python tools/ncu_one.py 14336 4096 4 64 1 gemv 6 > $O/${TAG}_ncu_gemv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm -s 1 -c 1 -f -o $O/${TAG}_gemm python tools/ncu_one.py 14336 4096 4 64 2048 gemm 3 > $O/${TAG}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny -s 2 -c 1 -f -o $O/${TAG}_skinny python tools/ncu_one.py 14336 4096 4 64 8 skinny 6 > $O/${TAG}_ncu_skinny.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
