#!/bin/bash
# r01e: recover measured state in one box: full GPU tests, chain floor / L2 prefetch probes, co-residency variants,
# grouped-launch bench, timelines.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r01e_pytest.log
tail -4 $O/r01e_pytest.log
bash tools/gpu_r01c.sh r01e_c
bash tools/gpu_r01d.sh r01e_d
