#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run6.txt
: > $O
( python tools/fh_debug.py generated_scene; python tools/fh_debug.py cornell_box; python tools/fh_debug.py synthetic_100000 ) 2>&1 | grep "variant=0 beam=1 smem_stack=1 stratify=1 smem_scene=1\|variant=4 beam=0 smem_stack=0" >> $O
python tools/exp.py generated_scene 0 4096 >> $O 2>&1
python tools/exp.py cornell_box 0 1024 >> $O 2>&1
python tools/exp_large.py 10000 256 >> $O 2>&1
python tools/exp_large.py 100000 256 >> $O 2>&1
cat $O
