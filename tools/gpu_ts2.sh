#!/bin/bash
# what bounds the TMEM-operand GEMM at small batches: bytes, row pieces per TMA box, or DRAM latency?
TAG=${1:-ts2}
O=gpurun_out; mkdir -p $O
{
timeout 300 python -m pytest tests/test_gpu_gather.py tests/test_gpu_parity.py -q -k "gather or transpose or edge or gemm_ts" 2>&1 | tail -12
echo "== base, all widths"; timeout 200 python tools/gemmbench.py --kernel gemm_ts --ms 16,64 --bits 2,3,4,6,8 2>&1 | grep -v "^3b"
for e in "GBXQ_TS_L2_PREFETCH=1" "GBXQ_TMA_L2PROMO=0" "GBXQ_TMA_L2PROMO=128" "GBXQ_TS_L2_PREFETCH=1 GBXQ_TMA_L2PROMO=0"; do
  echo "== $e"; env $e timeout 200 python tools/gemmbench.py --kernel gemm_ts --ms 16,64 --bits 2,4,8 2>&1 | grep -v "^3b\|^shape"
done
} > $O/${TAG}_bench.txt 2>&1
cat $O/${TAG}_bench.txt
