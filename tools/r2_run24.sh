#!/bin/bash
# round-2 GPU run 24 (2 GPUs): the multi-GPU context test through the C ABI, the smem-window test, bench.py at N = 2 and the binary's --gpus 2, on the final kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests -m gpu -q -s -k "multi_gpu or shared_memory_opt_in" 2>&1 | tail -6 | cut -c1-400 | tee gpurun_out/r2_run24_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02d_bench_ours_2gpu.json 2> gpurun_out/r02d_bench_ours_2gpu.err
tail -1 gpurun_out/r02d_bench_ours_2gpu.json | cut -c1-400
for g in 1 2; do
  ( cd assets && ../pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 4096 --gpus $g --stats scenes/generated_scene.json | tail -1 | sed "s/^{/{\"scene\": \"generated_scene 1920x1080\", \"spp\": 4096, \"gpus\": $g, /" ) | tee -a gpurun_out/r02d_cli_2gpu.jsonl | cut -c1-500
done
