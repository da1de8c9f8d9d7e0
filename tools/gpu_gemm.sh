#!/bin/bash
# Tensor-core GEMM work in one gpurun call: gemm_ts parity tests, per-shape sweep (TFLOP/s and GB/s), decode-batch and
# prefill bench lines.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_gemm.sh TAG'
TAG=${1:-ts3}
O=gpurun_out; mkdir -p $O
{
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "gemm_ts" 2>&1 | tail -4
echo "== rotate 1"; timeout 200 python tools/gemmbench.py --kernel gemm_ts --ms 8,16,32,64,128,2048 --bits 4,2 2>&1 | grep -v "^3b"
echo "== rotate 0"; GBXQ_TS_ROTATE=0 timeout 200 python tools/gemmbench.py --kernel gemm_ts --ms 16,2048 --bits 4 2>&1 | grep -v "^3b\|^shape"
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
for v in "--batch 8" "--batch 16" "--batch 32" "--batch 64" "--phase prefill --model llama-3.2-3b"; do
  echo "== TS_MIN_M=5 $v"; GBXQ_TS_MIN_M=5 timeout 200 python bench.py --no-cpu-baseline --steps 8 $v 2>&1 | tail -1 | b
done
} > $O/${TAG}_bench.txt 2>&1
cat $O/${TAG}_bench.txt
