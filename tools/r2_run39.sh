#!/bin/bash
# round-2 GPU run 39: A/B - node loads of global-memory scenes with the L1 evict_last hint
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run39.txt
: > $O
for v in head e1 head e1; do
  if [ $v = head ]; then export PT_B200_LIB=$PWD/build/exp/head/pathtracercuda_b200/libpt_b200.so; else export PT_B200_LIB=$PWD/build/exp/$v/libpt_b200.so; fi
  echo "== $v" >> $O
  for n in 10000 100000 1000000; do python tools/exp_large.py $n 256 2>&1 | tail -1 >> $O; done
done
unset PT_B200_LIB
grep -E "^==|\"ms\"" $O | sed -E 's/.*"objects": ([0-9]+).*"ms": ([0-9.]+).*/syn\1 \2/' | paste - - - - | head
