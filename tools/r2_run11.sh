#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run11.txt
: > $O
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/r2_run11_tests_full.txt
grep "first hits\|RMSE\|passed\|failed\|FAILED\|Error" gpurun_out/r2_run11_tests_full.txt | cut -c1-300 >> $O
python tools/exp.py generated_scene 0 4096 >> $O 2>&1
python tools/exp_large.py 10000 256 >> $O 2>&1
python tools/exp_large.py 100000 256 >> $O 2>&1
python tools/exp_large.py 1000000 256 >> $O 2>&1
cat $O
