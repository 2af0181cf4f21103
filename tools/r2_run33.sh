#!/bin/bash
# round-2 GPU run 33: env_is with deferred shadow rays (one-pixel-per-warp twin): tests and timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "env_importance or env_is" 2>&1 | tail -8 | cut -c1-600 | tee gpurun_out/r2_run33_envis_tests.txt
python tools/exp.py generated_scene 0 4096 env_is=1 2>&1 | tail -1 | cut -c1-300
python tools/exp.py generated_scene 0 256 env_is=1 2>&1 | tail -1 | cut -c1-300
python tools/exp.py generated_scene 0 4096 2>&1 | tail -1 | cut -c1-300
