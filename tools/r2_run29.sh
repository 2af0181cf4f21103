#!/bin/bash
# round-2 GPU run 29: the gates that see the 1 M-object scene, with the LBVH as its default builder
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -q -s -k "1000000" 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/r2_run29_tests.txt
