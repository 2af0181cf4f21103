#!/bin/bash
# round-2 GPU run 1: A/B timing of the kernel changes (HEAD copy under build/exp/head vs the working tree), then the GPU tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run1_exp.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv >> $O
echo "== HEAD (round 1 kernel)" >> $O
( cd build/exp/head && python tools/exp.py generated_scene 0 4096 && python tools/exp.py cornell_box 0 1024 ) >> $O 2>&1
echo "== tree: sort (round-1 algorithm) + FFMA2 + leaf class order, local stack" >> $O
python tools/exp.py generated_scene 0 4096 stratify=0 sort_samples=1 smem_stack=0 >> $O 2>&1
echo "== tree: + shared-memory stack" >> $O
python tools/exp.py generated_scene 0 4096 stratify=0 sort_samples=1 smem_stack=1 >> $O 2>&1
echo "== tree: no sort, no stratification, smem stack" >> $O
python tools/exp.py generated_scene 0 4096 stratify=0 sort_samples=0 smem_stack=1 >> $O 2>&1
echo "== tree: stratified, local stack" >> $O
python tools/exp.py generated_scene 0 4096 stratify=1 smem_stack=0 >> $O 2>&1
echo "== tree: default (stratified + smem stack)" >> $O
python tools/exp.py generated_scene 0 4096 >> $O 2>&1
python tools/exp.py generated_scene 0 4096 count_work=1 >> $O 2>&1
python tools/exp.py generated_scene 0 1024 >> $O 2>&1
python tools/exp.py generated_scene 0 256 >> $O 2>&1
echo "== cornell_box 1024" >> $O
python tools/exp.py cornell_box 0 1024 stratify=0 sort_samples=1 smem_stack=0 >> $O 2>&1
python tools/exp.py cornell_box 0 1024 stratify=0 sort_samples=1 smem_stack=1 >> $O 2>&1
python tools/exp.py cornell_box 0 1024 >> $O 2>&1
python tools/exp.py cornell_box 0 1024 count_work=1 >> $O 2>&1
echo "== synthetic 10k / 100k at 256 spp: HEAD, then tree (local stack / smem stack; stratify off / on)" >> $O
( cd build/exp/head && python tools/exp_large.py 10000 256 && python tools/exp_large.py 100000 256 ) >> $O 2>&1
python tools/exp_large.py 10000 256 stratify=0,smem_stack=0 stratify=0,smem_stack=1 stratify=1,smem_stack=1 >> $O 2>&1
python tools/exp_large.py 100000 256 stratify=0,smem_stack=0 stratify=0,smem_stack=1 stratify=1,smem_stack=1 >> $O 2>&1
cat $O
echo "== smoke" 
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2_run1_smoke.txt
echo "== tests"
timeout 2400 python -m pytest tests -m gpu -q -x -s --durations=15 2>&1 | tail -60 | tee gpurun_out/r2_run1_tests.txt
