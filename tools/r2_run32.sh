#!/bin/bash
# round-2 GPU run 32: __graft_entry__.smoke() on the committed tree
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -8 | tee gpurun_out/r2_run32_smoke.txt
