#!/bin/bash
# round-2 GPU run 2: FFMA2 issue microbenchmark, GPU tests, smoke, ncu capture of the default kernel (generated_scene 1024 spp)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
build/exp/ffma2_issue 2>&1 | tee gpurun_out/r2_ffma2_issue.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2_run2_smoke.txt
timeout 2400 python -m pytest tests -m gpu -q -s --durations=20 2>&1 | tail -120 | tee gpurun_out/r2_run2_tests.txt
python tools/prof_run.py 1024 > gpurun_out/r2_prof_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 1 -c 1 -o gpurun_out/r2a_trace_1024spp -f python tools/prof_run.py 1024 > gpurun_out/r2_ncu_log.txt 2>&1
ls -la gpurun_out/
