#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run5.txt
: > $O
echo "== first hits: tree" >> $O
( python tools/fh_debug.py generated_scene; python tools/fh_debug.py cornell_box; python tools/fh_debug.py synthetic_100000 ) 2>&1 | grep "variant=0 beam=1 smem_stack=1 stratify=1 smem_scene=1\|variant=4 beam=0 smem_stack=0" >> $O
echo "== first hits: exact build" >> $O
( export PT_B200_LIB=$PWD/build/exp/exact/libpt_b200.so; python tools/fh_debug.py generated_scene; python tools/fh_debug.py synthetic_100000 ) 2>&1 | grep "variant=0 beam=1 smem_stack=1 stratify=1 smem_scene=1\|variant=4 beam=0 smem_stack=0" >> $O
echo "== timing tree" >> $O
python tools/exp.py generated_scene 0 4096 >> $O 2>&1
python tools/exp.py cornell_box 0 1024 >> $O 2>&1
echo "== exact build timing" >> $O
PT_B200_LIB=$PWD/build/exp/exact/libpt_b200.so python tools/exp.py generated_scene 0 4096 >> $O 2>&1
echo "== regen_low sweep" >> $O
for r in 8 12 20 24 28; do python tools/exp.py generated_scene 0 4096 regen_low=$r 2>&1 | head -1 >> $O; done
echo "== strata_k sweep" >> $O
for k in 5 6 8 9; do python tools/exp.py generated_scene 0 4096 strata_k=$k 2>&1 | head -1 >> $O; done
cat $O
