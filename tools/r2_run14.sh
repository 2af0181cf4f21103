#!/bin/bash
# round-2 GPU run 14: A/B of micro-variants (separate builds under build/exp/*), three rounds interleaved so that box drift shows
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run14.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $O
for round in 1 2 3; do
for v in head base acos nosort surf all; do
  if [ $v = head ]; then export PT_B200_LIB=$PWD/build/exp/head/pathtracercuda_b200/libpt_b200.so; else export PT_B200_LIB=$PWD/build/exp/$v/libpt_b200.so; fi
  echo "== $v round $round" >> $O
  python tools/exp.py generated_scene 0 4096 2>&1 | head -1 >> $O
  if [ $round = 1 ]; then python tools/exp.py cornell_box 0 1024 2>&1 | head -1 >> $O; fi
done
done
export PT_B200_LIB=$PWD/build/exp/all/libpt_b200.so
for r in 12 20 24; do echo "== all regen_low=$r" >> $O; python tools/exp.py generated_scene 0 4096 regen_low=$r 2>&1 | head -1 >> $O; done
unset PT_B200_LIB
grep -E "^==|\"ms\"" $O | sed -E 's/.*"scene": "([a-z_]+)".*"ms": ([0-9.]+).*/\1 \2/' | paste - - | head -60
