#!/bin/bash
# round-2 GPU run 35: A/B - code that production never runs taken out of the hot kernel (software texture filter; max_bounces as a constant)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run35.txt
: > $O
for round in 1 2; do
for v in head t1 t2 t3; do
  if [ $v = head ]; then export PT_B200_LIB=$PWD/build/exp/head/pathtracercuda_b200/libpt_b200.so; else export PT_B200_LIB=$PWD/build/exp/$v/libpt_b200.so; fi
  echo "== $v round $round" >> $O
  python tools/exp.py generated_scene 0 4096 2>&1 | tail -1 >> $O
  python tools/exp.py cornell_box 0 1024 2>&1 | tail -1 >> $O
  if [ $round = 1 ]; then python tools/exp_large.py 10000 256 2>&1 | tail -1 >> $O; fi
done
done
unset PT_B200_LIB
grep -E "^==|\"ms\"|Error" $O | sed -E 's/.*"crc": ([0-9]+).*"scene": "([a-z_]+)".*"ms": ([0-9.]+).*/\2 \3 crc \1/; s/.*"objects": ([0-9]+).*"ms": ([0-9.]+).*/syn\1 \2/' | paste - - - - | head -40
