#!/usr/bin/env python3
"""Timing experiments on a synthetic scene: tools/exp_large.py N spp [key=value ...] (keys: pt_set_option keys)"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
n, spp = int(sys.argv[1]), int(sys.argv[2])
runs = [dict(a.split("=") for a in grp.split(",") if a) for grp in sys.argv[3:]] or [{}]
with tempfile.TemporaryDirectory() as td:
    os.symlink(pt.ASSETS + "/skybox.hdr", td + "/skybox.hdr")
    scenegen.write_synthetic_scene(td + "/scene.json", n)
    with pt.Pathtracer(1920, 1080) as P:
        cam = P.loadSceneFile(td + "/scene.json", cwd=td)
        P.render(cam, 8, True)
        for kv in runs:
            for k, v in kv.items():
                P.setOption(k, float(v))
            best = min((P.render(cam, spp, True), P.getTiming())[1] for _ in range(3))
            st = P.stats()
            print(json.dumps({"objects": n + 1, "spp": spp, "opts": kv, "ms": round(best, 3), "Mrays_s": round(st.rays / best / 1e3, 1)}), flush=True)
