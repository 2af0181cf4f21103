#!/bin/bash
# round-2 GPU run 25: ncu full capture of the finalised trace kernel (generated_scene 1080p, 1024 spp) + launch list of the bench command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/prof_run.py 1024 > gpurun_out/r2d_prof_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 1 -c 1 -o gpurun_out/r2d_trace_1024spp -f python tools/prof_run.py 1024 > gpurun_out/r2d_ncu_log.txt 2>&1
cat gpurun_out/r2d_prof_plain.txt
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-cli > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-cli > gpurun_out/r2d_ncu_launch_log.txt 2>&1
tail -5 gpurun_out/r02d_ncu_launches_bench.csv | cut -c1-200
ls -la gpurun_out/r2d_trace_1024spp.ncu-rep
