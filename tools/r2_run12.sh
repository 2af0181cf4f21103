#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -3 gpurun_out/r2_bench_ref.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_ours.json 2> gpurun_out/r2_bench_ours.err; tail -3 gpurun_out/r2_bench_ours.err
python bench.py --scene cornell_box --spp 1024 --steps 3 --warmup 3 > gpurun_out/r2_bench_ours_cornell.json 2> gpurun_out/r2_bench_ours_cornell.err; tail -3 gpurun_out/r2_bench_ours_cornell.err
python bench.py --impl reference --scene cornell_box --spp 1024 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_cornell.json 2>> gpurun_out/r2_bench_ref.err
python bench.py --scene synthetic_100000 --spp 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ours_syn100k.json 2> gpurun_out/r2_bench_ours_syn100k.err; tail -3 gpurun_out/r2_bench_ours_syn100k.err
cat gpurun_out/r2_bench_ref.json gpurun_out/r2_bench_ours.json | cut -c1-1500
