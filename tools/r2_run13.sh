#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/r2_run13_tests_full.txt
grep "first hits\|RMSE\|passed\|failed\|FAILED\|Error" gpurun_out/r2_run13_tests_full.txt | cut -c1-300
( cd assets && for i in 1 2; do ../pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 4096 -ohdr -o /tmp/o.hdr --stats scenes/generated_scene.json | tail -1; done ) | tee gpurun_out/r2_run13_cli.txt
# launch list of the bench (device time per kernel), then full captures of the trace kernel on four more workloads
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-cli > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e-cli > gpurun_out/r2_ncu_launch_log.txt 2>&1
for job in "1024 0 cornell_box r02_cornell_1024spp" "256 0 synthetic_10000 r02_syn10k_256spp" "256 0 synthetic_100000 r02_syn100k_256spp" "256 0 synthetic_1000000 r02_syn1m_256spp"; do
  set -- $job
  python tools/prof_run.py $1 $2 $3 > gpurun_out/$4.plain.txt 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 1 -c 1 -o gpurun_out/$4 -f python tools/prof_run.py $1 $2 $3 > gpurun_out/$4.ncu.log 2>&1
  cat gpurun_out/$4.plain.txt | tail -1
done
ls -la gpurun_out/*.ncu-rep | tail -6
