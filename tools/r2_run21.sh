#!/bin/bash
# round-2 GPU run 21: A/B (material table in shared memory, pinned addresses) + host load phases of the 1 M-object scene with the mapped file / single-scan splitter
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run21.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $O
nproc >> $O
for round in 1 2; do
for v in head m0 m1 m2 m3 m4 tree; do
  if [ $v = head ]; then export PT_B200_LIB=$PWD/build/exp/head/pathtracercuda_b200/libpt_b200.so; elif [ $v = tree ]; then unset PT_B200_LIB; else export PT_B200_LIB=$PWD/build/exp/$v/libpt_b200.so; fi
  echo "== $v round $round" >> $O
  python tools/exp.py generated_scene 0 4096 2>&1 | tail -1 >> $O
  if [ $round = 1 ]; then python tools/exp.py cornell_box 0 1024 2>&1 | tail -1 >> $O; fi
done
done
unset PT_B200_LIB
grep -E "^==|\"ms\"|Error" $O | sed -E 's/.*"crc": ([0-9]+).*"scene": "([a-z_]+)".*"ms": ([0-9.]+).*/\2 \3 crc \1/' | paste - - - | head -60
python -c "
import sys; sys.path.insert(0,'.')
from pathtracercuda_b200 import scenegen
import os
os.symlink(os.path.abspath('assets/skybox.hdr'), '/tmp/skybox.hdr')
scenegen.write_synthetic_scene('/tmp/syn1m.json', 1000000)
"
( cd /tmp && for i in 1 2 3; do PTB_TIMING=1 $GRAFT_REPO_ROOT/pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 256 -ohdr -o /tmp/o.hdr --stats /tmp/syn1m.json 2>&1 | grep -E "compileScene|loadScene|host_ms" | sed -E 's/.*("host_ms": \{[^}]*\}).*/\1/'; done ) | tee gpurun_out/r2_run21_load1m.txt
