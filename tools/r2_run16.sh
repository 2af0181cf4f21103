#!/bin/bash
# round-2 GPU run 16: A/B of micro-variants, second batch (the three changes of run 14's "base" apart; packed primitive records)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run16.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $O
for round in 1 2; do
for v in head b0 b1 b2 b3 b4 b5 b6 b7; do
  if [ $v = head ]; then export PT_B200_LIB=$PWD/build/exp/head/pathtracercuda_b200/libpt_b200.so; else export PT_B200_LIB=$PWD/build/exp/$v/libpt_b200.so; fi
  echo "== $v round $round" >> $O
  python tools/exp.py generated_scene 0 4096 2>&1 | head -1 >> $O
  if [ $round = 1 ]; then python tools/exp.py cornell_box 0 1024 2>&1 | head -1 >> $O; fi
done
done
unset PT_B200_LIB
grep -E "^==|\"ms\"" $O | sed -E 's/.*"crc": ([0-9]+).*"scene": "([a-z_]+)".*"ms": ([0-9.]+).*/\2 \3 crc \1/' | paste - - - | head -60
