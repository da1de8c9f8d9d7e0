#!/bin/bash
# A/B of two builds of the library on the default bench, variants and per-shape sweeps.
# Usage: gpurun -- 'bash tools/gpu_ab.sh TAG libA.so libB.so'   (paths relative to gbx-lm_b200/)
TAG=${1:-ab}; A=$2; B=$3
O=gpurun_out; mkdir -p $O
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"; }
{
for lib in $A $B; do
  export GBXQ_LIB=$PWD/gbx-lm_b200/$lib
  echo "#### $lib"
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "mmv8 or grouped" 2>&1 | tail -2
  for v in "" "--strategy bpw-2.2" "--batch 2" "--model llama-3-70b --steps 5" "--model llama-3.2-3b"; do
    echo "== $v"; timeout 200 python bench.py --no-cpu-baseline $v 2>&1 | tail -1 | b
  done
  timeout 200 python tools/microbench.py --quick --kernel mmv8 --ms 1 2>&1
done
} > $O/${TAG}_ab.txt 2>&1
cat $O/${TAG}_ab.txt
