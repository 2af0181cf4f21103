#!/bin/bash
# round-2 GPU run 40: regen_low on the global-memory scenes (the default 16 was tuned on generated_scene)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for n in 10000 1000000; do python tools/exp_large.py $n 256 regen_low=16 regen_low=8 regen_low=12 regen_low=20 regen_low=24 regen_low=28 2>&1 | cut -c1-200; done | tee gpurun_out/r2_run40.txt
