#!/bin/bash
# round-2 GPU run 26: compute-sanitizer (memcheck, racecheck, synccheck) over small renders of every default code path and the twin instantiations;
# the full GPU test suite on the last host-side changes; load phases of the 1 M-object scene
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/sanitize_run.py > gpurun_out/r02d_sanitize_plain.txt 2>&1; tail -3 gpurun_out/r02d_sanitize_plain.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/r02d_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" gpurun_out/r02d_sanitize_$tool.txt | head -5
done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_run26_tests.txt
python -c "
import sys; sys.path.insert(0,'.')
from pathtracercuda_b200 import scenegen
import os
os.symlink(os.path.abspath('assets/skybox.hdr'), '/tmp/skybox.hdr')
scenegen.write_synthetic_scene('/tmp/syn1m.json', 1000000)
"
( cd /tmp && for i in 1 2 3; do PTB_TIMING=1 $GRAFT_REPO_ROOT/pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 256 -ohdr -o /tmp/o.hdr --stats /tmp/syn1m.json 2>&1 | grep -E "compileScene|loadScene|host_ms" | sed -E 's/.*("host_ms": \{[^}]*\}).*/\1/'; done ) | tee gpurun_out/r2_run26_load1m.txt | grep host_ms
