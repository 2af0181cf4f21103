#!/usr/bin/env python3
"""Short run of the trace kernel for ncu: tools/prof_run.py [spp] [variant] [scene | synthetic_<N>] [W H]  (3 launches)"""
import os
import sys
import tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scene = sys.argv[3] if len(sys.argv) > 3 else "generated_scene"
W, H = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (1920, 1080)
with tempfile.TemporaryDirectory() as td, pt.Pathtracer(W, H) as P:
    if scene.startswith("synthetic_"):
        os.symlink(pt.ASSETS + "/skybox.hdr", td + "/skybox.hdr")
        scenegen.write_synthetic_scene(td + "/scene.json", int(scene.split("_")[1]))
        cam = P.loadSceneFile(td + "/scene.json", cwd=td)
    else:
        cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/{scene}.json", cwd=pt.ASSETS)
    P.setOption("variant", variant)
    for i in range(3):
        P.render(cam, spp, True)
        st = P.stats()
        print(f"{scene} {W}x{H} spp={spp} variant={variant}: {P.getTiming():.3f} ms  {st.rays / P.getTiming() / 1e3:.1f} Mrays/s")
