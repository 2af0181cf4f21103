#!/bin/bash
# round-2 GPU run 30: the finalised kernel (packed primitive records, inlined miss path, acos polynomial, no sample sort):
# full GPU test suite, both bench arms, cornell_box + large scenes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/r2_run30.txt
: > $O
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/r2_run30_tests_full.txt
grep "first hits\|RMSE\|passed\|failed\|FAILED\|Error" gpurun_out/r2_run30_tests_full.txt | cut -c1-300 >> $O
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench_ours.json 2> gpurun_out/r2e_bench_ours.err
python bench.py --scene cornell_box --spp 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-cli > gpurun_out/r2e_bench_ours_cornell.json 2>> $O
python tools/exp.py generated_scene 0 4096 >> $O 2>&1
python tools/exp_large.py 10000 256 >> $O 2>&1
python tools/exp_large.py 100000 256 >> $O 2>&1
python tools/exp_large.py 1000000 256 >> $O 2>&1
cat $O
python -c "
import json
for f in ('r2e_bench_ref','r2e_bench_ours','r2e_bench_ours_cornell'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('e2e'), d.get('roofline',{}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
"
