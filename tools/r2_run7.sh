#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/r2_run7_tests_full.txt
tail -15 gpurun_out/r2_run7_tests_full.txt
python tools/prof_run.py 1024 > gpurun_out/r2_prof_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 1 -c 1 -o gpurun_out/r2b_trace_1024spp -f python tools/prof_run.py 1024 > gpurun_out/r2_ncu_log.txt 2>&1
cat gpurun_out/r2_prof_plain.txt
