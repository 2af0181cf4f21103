#!/bin/bash
# round-2 multi-GPU measurements on one 8-GPU box: config 3 through bench.py (torchrun, one process per GPU) at N = 2, 4, 8;
# config 4 (synthetic 10 k / 100 k / 1 M, 256 spp) and config 5 (4K, 16384 spp) through the drop-in binary's --gpus N
# (C ABI pt_create_multi: one process, one host thread per GPU) at N = 1, 2, 4, 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
python - <<'PY'
import os, time
from pathtracercuda_b200 import scenegen
import pathtracercuda_b200 as pt
os.makedirs("/tmp/ptb_scenes", exist_ok=True)
if not os.path.exists("/tmp/ptb_scenes/skybox.hdr"):
    os.symlink(pt.ASSETS + "/skybox.hdr", "/tmp/ptb_scenes/skybox.hdr")
for n in (10000, 100000, 1000000):
    t = time.time(); scenegen.write_synthetic_scene(f"/tmp/ptb_scenes/synthetic_{n}.json", n); print("generated", n, round(time.time() - t, 1), "s", flush=True)
PY
O=gpurun_out/r02_config4_config5_cli.jsonl
: > $O
for n in 10000 100000 1000000; do
  for g in 1 2 4 8; do
    for rep in 1 2; do
      ( cd /tmp/ptb_scenes && $OLDPWD/pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 256 --gpus $g --stats synthetic_$n.json | tail -1 | sed "s/^{/{\"scene\": \"synthetic_$n\", \"spp\": 256, \"gpus\": $g, /" ) >> $O
    done
  done
done
for ex in p2p nccl; do
  ( cd assets && ../pathtracercuda_b200/bin/pathtracer_b200 -w 3840 -h 2160 -spp 16384 --gpus 8 --exchange $ex --stats scenes/generated_scene.json | tail -1 | sed "s/^{/{\"scene\": \"generated_scene 3840x2160\", \"spp\": 16384, \"gpus\": 8, /" ) >> $O
done
( cd assets && ../pathtracercuda_b200/bin/pathtracer_b200 -w 3840 -h 2160 -spp 16384 --gpus 8 --partition samples --stats scenes/generated_scene.json | tail -1 | sed "s/^{/{\"scene\": \"generated_scene 3840x2160\", \"spp\": 16384, \"gpus\": 8, /" ) >> $O
for g in 1 2 4 8; do
  ( cd assets && ../pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 4096 --gpus $g --stats scenes/generated_scene.json | tail -1 | sed "s/^{/{\"scene\": \"generated_scene 1920x1080\", \"spp\": 4096, \"gpus\": $g, /" ) >> $O
done
cat $O | cut -c1-400
for g in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2950$g bench.py --gpus $g --steps 3 --warmup 3 > gpurun_out/r02_bench_ours_${g}gpu.json 2> gpurun_out/r02_bench_ours_${g}gpu.err
  tail -1 gpurun_out/r02_bench_ours_${g}gpu.json | cut -c1-300
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --partition samples --no-e2e-cli > gpurun_out/r02_bench_ours_8gpu_samples.json 2> gpurun_out/r02_bench_ours_8gpu_samples.err
tail -1 gpurun_out/r02_bench_ours_8gpu_samples.json | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "multi_gpu" 2>&1 | tail -3
