#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( python tools/fh_debug.py generated_scene; python tools/fh_debug.py cornell_box; python tools/fh_debug.py synthetic_100000 ) 2>&1 | tee gpurun_out/r2_run4_fh.txt
