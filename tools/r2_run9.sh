#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run9.txt
: > $O
python tools/exp.py generated_scene 0 4096 >> $O 2>&1
python tools/exp.py cornell_box 0 1024 >> $O 2>&1
python tools/exp_large.py 10000 256 >> $O 2>&1
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -q -s -k "first_hit" 2>&1 | grep "first hits\|passed\|failed\|Error" >> $O
cat $O
