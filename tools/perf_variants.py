#!/usr/bin/env python3
"""Development probe: time kernel variants on generated_scene / cornell_box 1080p and check they agree bit for bit."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracercuda_b200 as pt

variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 4, 5]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
scenes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["generated_scene", "cornell_box"]
extra = dict(kv.split("=") for kv in sys.argv[4].split(",")) if len(sys.argv) > 4 else {}
ref_img = {}
for scene in scenes:
    for v in variants:
        with pt.Pathtracer(1920, 1080) as P:
            cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/{scene}.json", cwd=pt.ASSETS)
            P.setOption("variant", v)
            for k, val in extra.items():
                P.setOption(k, float(val))
            P.render(cam, 8, True)
            best = 1e30
            for _ in range(3):
                P.render(cam, spp, True)
                best = min(best, P.getTiming())
            st = P.stats()
            img = P.getHDRMean()
            if scene not in ref_img:
                ref_img[scene] = img
            same = bool(np.array_equal(ref_img[scene].view(np.uint32), img.view(np.uint32)))
            print(json.dumps({"scene": scene, "variant": v, "ms": round(best, 3), "Mrays_s": round(st.rays / best / 1e3, 1), "Msamples_s": round(st.samples / best / 1e3, 1),
                              "extra": extra, "bvh_nodes": st.bvh_nodes, "nodes_per_ray": round(st.node_visits / max(st.rays, 1), 2), "prims_per_ray": round(st.prim_tests / max(st.rays, 1), 2), "bit_identical_to_first": same, "maxdiff": float(np.abs(ref_img[scene] - img).max())}), flush=True)
