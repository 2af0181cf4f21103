#!/bin/bash
# round-2 GPU run 38: cornell_box with fewer hoisted primitives (option max_global), generated_scene as the control
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for g in 8 6 4 2 0; do
  python tools/exp.py cornell_box 0 1024 max_global=$g count_work=1 2>&1 | head -1 | cut -c1-330
  python tools/exp.py cornell_box 0 1024 max_global=$g 2>&1 | head -1 | cut -c1-200
done | tee gpurun_out/r2_run38.txt
python tools/exp.py generated_scene 0 4096 max_global=0 2>&1 | head -1 | cut -c1-200 | tee -a gpurun_out/r2_run38.txt
