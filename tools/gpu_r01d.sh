#!/bin/bash
# r01d: grouped launch parity + bench, L2-resident (compute ceiling) sweep, finer prologue timeline.
TAG=${1:-r01d}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped or mmv or cuda_graph or auto_dispatch" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; }
{
for v in "--grouped 1" "--grouped 0" "--grouped 1 --strategy bpw-2.2" "--grouped 1 --batch 4"; do
  echo "== $v"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v | b
done
echo "== L2-resident weights (one copy): compute / L2 ceiling of the kernel"
timeout 300 python tools/microbench.py --quick --kernel mmv8 --ms 1 --l2 2>&1
GBXQ_MMV8_GRID_MULT=1 timeout 300 python tools/microbench.py --quick --kernel mmv8 --ms 1 --l2 2>&1 | grep -v shape
for shp in "14336 4096 4 64" "4096 14336 4 64" "1024 4096 4 64"; do
  echo "== timeline $shp"; timeout 120 python tools/timeline.py $shp 8 2
done
} > $O/${TAG}_bench.txt 2>&1
cat $O/${TAG}_bench.txt
