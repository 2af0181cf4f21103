#!/bin/bash
# round-2 GPU run 28: the LBVH builder (option bvh_builder = 1): GPU test, render and load times against the SAH builder on 10 k / 100 k / 1 M objects, the first-hit gate at 1 M with it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "lbvh" 2>&1 | tail -6 | cut -c1-400 | tee gpurun_out/r2_run28_tests.txt
python - <<'PY' | tee gpurun_out/r02d_lbvh_vs_sah.jsonl
import json, os, sys, tempfile, time
sys.path.insert(0, ".")
import numpy as np
import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
for n in (10000, 100000, 1000000):
    with tempfile.TemporaryDirectory() as td:
        os.symlink(pt.ASSETS + "/skybox.hdr", td + "/skybox.hdr")
        scenegen.write_synthetic_scene(td + "/scene.json", n)
        for builder in (0, 1):
            with pt.Pathtracer(1920, 1080) as P:
                P.setOption("bvh_builder", builder)
                t0 = time.perf_counter(); cam = P.loadSceneFile(td + "/scene.json", cwd=td); load = time.perf_counter() - t0
                t0 = time.perf_counter(); cam = P.loadSceneFile(td + "/scene.json", cwd=td); load2 = time.perf_counter() - t0
                P.render(cam, 8, True)
                best = min((P.render(cam, 256, True), P.getTiming())[1] for _ in range(3))
                st = P.stats()
                print(json.dumps({"objects": n + 1, "builder": ["sah", "lbvh"][builder], "load_s": round(min(load, load2), 3), "render_ms_256spp": round(best, 3), "Mrays_s": round(st.rays / best / 1e3, 1),
                                  "bvh_nodes": st.bvh_nodes, "bvh_depth": st.bvh_depth, "rays_per_sample": round(st.rays / st.samples, 4)}), flush=True)
PY
( cd /tmp && python -c "
import sys; sys.path.insert(0,'$GRAFT_REPO_ROOT')
from pathtracercuda_b200 import scenegen
import os
os.symlink('$GRAFT_REPO_ROOT/assets/skybox.hdr', '/tmp/skybox.hdr')
scenegen.write_synthetic_scene('/tmp/syn1m.json', 1000000)
" && for b in sah lbvh; do PTB_TIMING=1 $GRAFT_REPO_ROOT/pathtracercuda_b200/bin/pathtracer_b200 -w 1920 -h 1080 -spp 256 -ohdr -o /tmp/o.hdr --bvh $b --stats /tmp/syn1m.json 2>&1 | grep -E "compileScene build|host_ms" | sed -E 's/.*("host_ms": \{[^}]*\}).*/\1/'; done ) | tee gpurun_out/r2_run28_load.txt
