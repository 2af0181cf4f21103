#!/bin/bash
# round-2 GPU run 22: what each parity aid costs the hot loop (separate builds), env_is timing with any-hit shadow rays
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run22.txt
: > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $O
for round in 1 2; do
for v in head a0 a1 a2 a3 a4; do
  if [ $v = head ]; then export PT_B200_LIB=$PWD/build/exp/head/pathtracercuda_b200/libpt_b200.so; else export PT_B200_LIB=$PWD/build/exp/$v/libpt_b200.so; fi
  echo "== $v round $round" >> $O
  python tools/exp.py generated_scene 0 4096 2>&1 | tail -1 >> $O
  if [ $round = 1 ]; then python tools/exp.py cornell_box 0 1024 2>&1 | tail -1 >> $O; fi
done
done
export PT_B200_LIB=$PWD/build/exp/a0/libpt_b200.so
echo "== a0 env_is=1 4096" >> $O; python tools/exp.py generated_scene 0 4096 env_is=1 2>&1 | tail -1 >> $O
echo "== a0 env_is=1 256" >> $O; python tools/exp.py generated_scene 0 256 env_is=1 2>&1 | tail -1 >> $O
echo "== a0 env_is=0 256" >> $O; python tools/exp.py generated_scene 0 256 2>&1 | tail -1 >> $O
unset PT_B200_LIB
grep -E "^==|\"ms\"|Error" $O | sed -E 's/.*"crc": ([0-9]+).*"scene": "([a-z_]+)".*"ms": ([0-9.]+).*/\2 \3 crc \1/' | paste - - - | head -60
