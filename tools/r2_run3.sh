#!/bin/bash
# round-2 GPU run 3: thread-count / inlining variants (A/B on one box), then the failing parity tests with full output
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run3_exp.txt
: > $O
for v in tree t896 t768 inl inl896; do
  if [ $v = tree ]; then unset PT_B200_LIB; else export PT_B200_LIB=$PWD/build/exp/$v/libpt_b200.so; fi
  echo "== $v" >> $O
  python tools/exp.py generated_scene 0 4096 >> $O 2>&1
  python tools/exp.py cornell_box 0 1024 >> $O 2>&1
  python tools/exp_large.py 10000 256 >> $O 2>&1
done
unset PT_B200_LIB
echo "== tree, local stack / no stratification (what each is worth now)" >> $O
python tools/exp.py generated_scene 0 4096 smem_stack=0 >> $O 2>&1
python tools/exp.py generated_scene 0 4096 stratify=0 >> $O 2>&1
cat $O
timeout 1200 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -q -s 2>&1 > gpurun_out/r2_run3_tests_full.txt
tail -5 gpurun_out/r2_run3_tests_full.txt
