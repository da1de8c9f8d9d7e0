#!/bin/bash
# ncu evidence of the round: launch list of one default bench step (time + DRAM bytes per launch), and --set full
# captures of the decode kernel, the TMEM-operand GEMM on a prefill tile and on a decode batch.
TAG=${1:-ncu}
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mmv8 -c 400 --csv \
  --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
tail -2 $O/${TAG}_launches.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmv8_grouped -s 4 -c 1 -f -o $O/${TAG}_mmv8_grouped \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu1.log 2>&1; tail -1 $O/${TAG}_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ts -s 3 -c 1 -f -o $O/${TAG}_ts_prefill \
  python tools/ncu_one.py 14336 4096 4 64 2048 gemm_ts > $O/${TAG}_ncu2.log 2>&1; tail -1 $O/${TAG}_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ts -s 3 -c 1 -f -o $O/${TAG}_ts_m32 \
  python tools/ncu_one.py 14336 4096 4 64 32 gemm_ts > $O/${TAG}_ncu3.log 2>&1; tail -1 $O/${TAG}_ncu3.log
ls -la $O/${TAG}_*
