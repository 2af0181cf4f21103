#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/fh_debug2.py 1000000 2>&1 | tee gpurun_out/r2_run10.txt
