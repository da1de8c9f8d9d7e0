#!/bin/bash
# TMEM-operand GEMM: the GPU suite, then per-shape sweeps old vs new and the decode-batch / prefill bench lines.
TAG=${1:-ts}
O=gpurun_out; mkdir -p $O
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
timeout 700 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest.log; tail -8 $O/${TAG}_pytest.log
{
for k in gemm_ts; do echo "== $k"; timeout 200 python tools/gemmbench.py --kernel $k --ms 8,16,32,64,128,2048 --bits 4,2 2>&1 | grep -v "^3b"; done
for v in "" "--batch 8" "--batch 16" "--batch 32" "--batch 64" "--phase prefill --model llama-3.2-3b"; do
  for ts in 0 5; do echo "== TS_MIN_M=$ts $v"; GBXQ_TS_MIN_M=$ts timeout 200 python bench.py --no-cpu-baseline --steps 8 $v 2>&1 | tail -1 | b; done
done
} > $O/${TAG}_bench.txt 2>&1
cat $O/${TAG}_bench.txt
