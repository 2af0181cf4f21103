#!/usr/bin/env python3
"""Short single-launch run of the trace kernel for ncu (generated_scene 1080p, few spp)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracercuda_b200 as pt
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
scene = sys.argv[2] if len(sys.argv) > 2 else "generated_scene"
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
with pt.Pathtracer(W, H) as P:
    cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/{scene}.json", cwd=pt.ASSETS)
    for i in range(3):
        P.render(cam, spp, True)
        st = P.stats()
        print(f"{scene} {W}x{H} spp={spp}: {P.getTiming():.3f} ms  {st.rays / P.getTiming() / 1e3:.1f} Mrays/s")
