#!/bin/bash
# ring geometry of the 16-warp decode kernel: stages in the ring, pre-issued stages, stage / ring size (bench.py per setting)
b() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'])"; }
for e in "GBXQ_MMV8_STAGES=3 GBXQ_MMV8_PRE_STAGES=3" "GBXQ_MMV8_STAGES=2" "GBXQ_MMV8_STAGES=3 GBXQ_MMV8_STAGE_KB=72 GBXQ_MMV8_RING_KB=216" "GBXQ_MMV8_STAGES=3 GBXQ_MMV8_STAGE_KB=48 GBXQ_MMV8_RING_KB=144"; do
  echo "== [$e]"; env $e timeout 200 python bench.py --no-cpu-baseline --no-also 2>&1 | tail -1 | b
done
for v in "--model llama-3-70b --steps 5" "--model llama-3.2-3b" "--strategy bpw-2.2"; do
  for e in "X=1" "GBXQ_MMV8_STAGES=3"; do echo "== [$e] $v"; env $e timeout 200 python bench.py --no-cpu-baseline --no-also $v 2>&1 | tail -1 | b; done
done
