#!/bin/bash
# round-2 GPU run 34: ncu full captures of the FINAL trace kernel on cornell_box (config 2) and the 100 k / 1 M-object scenes (config 4)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for job in "1024 0 cornell_box r02e_cornell_1024spp" "256 0 synthetic_100000 r02e_syn100k_256spp" "256 0 synthetic_1000000 r02e_syn1m_256spp"; do
  set -- $job
  python tools/prof_run.py $1 $2 $3 > gpurun_out/$4.plain.txt 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 1 -c 1 -o gpurun_out/$4 -f python tools/prof_run.py $1 $2 $3 > gpurun_out/$4.ncu.log 2>&1
  tail -1 gpurun_out/$4.plain.txt
done
ls -la gpurun_out/r02e_*.ncu-rep
