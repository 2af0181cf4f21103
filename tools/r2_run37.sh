#!/bin/bash
# round-2 GPU run 37: last check of the committed tree: smoke(), the full GPU test suite, the default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300
python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; tail -1 gpurun_out/r2f_bench_default.json | cut -c1-400; tail -3 gpurun_out/r2f_bench_default.err
