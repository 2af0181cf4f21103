#!/usr/bin/env python3
"""Short single-launch run of the trace kernel for ncu (generated_scene 1080p, few spp)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracercuda_b200 as pt
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scene = sys.argv[3] if len(sys.argv) > 3 else "generated_scene"
W, H = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (1920, 1080)
with pt.Pathtracer(W, H) as P:
    cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/{scene}.json", cwd=pt.ASSETS)
    P.setOption("variant", variant)
    for i in range(3):
        P.render(cam, spp, True)
        st = P.stats()
        print(f"{scene} {W}x{H} spp={spp} variant={variant}: {P.getTiming():.3f} ms  {st.rays / P.getTiming() / 1e3:.1f} Mrays/s")
