#!/bin/bash
# The device-side quantiser in one short call: its GPU tests, the microbench, an ncu launch list with DRAM bytes and
# (FULL=1) the default bench line of the library.  Usage: gpurun --timeout 200 -- 'bash tools/gpu_quant.sh TAG'
TAG=${1:-quant}
O=gpurun_out; mkdir -p $O
timeout 90 python -m pytest tests/test_gpu_quantize.py -m gpu -x -q > $O/${TAG}_pytest_quantize.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_quantize.log; tail -15 $O/${TAG}_pytest_quantize.log
timeout 40 python tools/quantbench.py > $O/${TAG}_quantbench.txt 2>&1; cat $O/${TAG}_quantbench.txt
if [ -n "$FULL" ]; then timeout 60 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err; fi
timeout 50 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:quantize -c 75 --csv --log-file $O/${TAG}_ncu_quantize.csv python tools/quantbench.py 'weight matrix' > /dev/null 2>&1; grep -c quantize_kernel $O/${TAG}_ncu_quantize.csv
