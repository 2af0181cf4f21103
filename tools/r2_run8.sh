#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "multi_gpu" 2>&1 | tail -40 | tee gpurun_out/r2_run8_multi.txt
cd assets
for ex in p2p nccl; do
  ../pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 4096 --gpus 2 --exchange $ex --stats scenes/generated_scene.json 2>&1 | tail -2 | tee -a ../gpurun_out/r2_run8_multi.txt
done
../pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 4096 --gpus 2 --partition samples --stats scenes/generated_scene.json 2>&1 | tail -2 | tee -a ../gpurun_out/r2_run8_multi.txt
../pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 4096 --stats scenes/generated_scene.json 2>&1 | tail -2 | tee -a ../gpurun_out/r2_run8_multi.txt
