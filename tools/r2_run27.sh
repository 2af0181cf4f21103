#!/bin/bash
# round-2 GPU run 27 (8 GPUs): the final kernel at N = 8 - bench.py (config 3, pixel partition) and the binary on config 5 (4K, 16384 spp; p2p and nccl exchange)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e-cli > gpurun_out/r02d_bench_ours_8gpu.json 2> gpurun_out/r02d_bench_ours_8gpu.err
tail -1 gpurun_out/r02d_bench_ours_8gpu.json | cut -c1-300
O=gpurun_out/r02d_config5_cli.jsonl
: > $O
for ex in p2p nccl; do
  ( cd assets && ../pathtracercuda_b200/bin/pathtracer_b200 -w 3840 -h 2160 -spp 16384 --gpus 8 --exchange $ex --stats scenes/generated_scene.json | tail -1 | sed "s/^{/{\"scene\": \"generated_scene 3840x2160\", \"spp\": 16384, \"gpus\": 8, /" ) >> $O
done
( cd assets && ../pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 4096 --gpus 8 --stats scenes/generated_scene.json | tail -1 | sed "s/^{/{\"scene\": \"generated_scene 1920x1080\", \"spp\": 4096, \"gpus\": 8, /" ) >> $O
cat $O | cut -c1-420
