#!/usr/bin/env python3
"""sort_samples on/off: same rays, same image up to float summation order, deterministic"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pathtracercuda_b200 as pt
imgs = []
for so in (0, 1, 1):
    with pt.Pathtracer(320, 180) as P:
        cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
        P.setOption("variant", 12); P.setOption("sort_samples", so)
        P.render(cam, 1024, True)
        imgs.append(P.getHDRMean()); print(so, P.stats().rays)
print("sorted vs unsorted allclose", np.allclose(imgs[0], imgs[1], rtol=3e-5, atol=1e-7), float(np.abs(imgs[0]-imgs[1]).max()), "deterministic", np.array_equal(imgs[1], imgs[2]))
