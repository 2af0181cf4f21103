#!/usr/bin/env python3
"""GPU-side golden fixtures, produced by the REFERENCE's own device code on a B200 (run under gpurun; copy
gpurun_out/golden_gpu/*.npz to tests/golden/):
  refgpu_primary_<scene>.npz   scene-order hit index + t of pixel-centre rays through the reference hitBVH
                               (oracle/_ref/ref_gpu = oracle/ref_harness/ref_gpu.cu around the reference trace.cu)
  refpt_<scene>.npz            4096-spp linear HDR renders of the UNMODIFIED reference program (oracle/_ref/ref_pt,
                               seed base 1984) and of its independent-seed twin (ref_pt_seedB, base 7919), decoded
                               from the .hdr files it writes and divided by 8 to undo its headless normalisation (Q1),
                               plus the reference's ray count for that configuration (ref_gpu count)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
from oracle import imgio, orc

OUT = os.path.join(ROOT, "gpurun_out", "golden_gpu")
os.makedirs(OUT, exist_ok=True)

with tempfile.TemporaryDirectory() as td:
    for scene, (W, H) in {"cornell_box": (256, 256), "generated_scene": (480, 270), "synthetic_1500": (320, 180)}.items():
        if scene.startswith("synthetic"):
            objs, cam = scenegen.synthetic_scene(1500, W, H)
        else:
            objs, tex, sky, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{scene}.json", W, H)
        idx, t = orc.ref_gpu_primary(objs, cam, W, H, td)
        np.savez_compressed(os.path.join(OUT, f"refgpu_primary_{scene}.npz"), W=W, H=H, idx=idx, t=t)
        print(scene, "primary", W, H, float((idx >= 0).mean()), flush=True)

    for scene, (W, H, spp) in {"cornell_box": (256, 256, 4096), "generated_scene": (320, 180, 4096)}.items():
        imgs = {}
        for tag, exe in (("A", orc.REF_PT), ("B", orc.REF_PT_SEEDB)):
            out = os.path.join(td, f"{scene}_{tag}.hdr")
            p = subprocess.run([exe, "-w", str(W), "-h", str(H), "-spp", str(spp), "-ohdr", "-o", out, f"scenes/{scene}.json"], cwd=pt.ASSETS,
                               capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            ms = [l for l in p.stdout.splitlines() if l.startswith("Finished")][0]
            img = imgio.read_hdr(out)[::-1, :, :3]  # file is top-down; buffers are bottom-up
            imgs[tag] = (img / np.float32(spp / ((spp + 7) // 8))).astype(np.float32)  # undo sum / ceil(spp/8)
            print(scene, tag, ms, "mean", float(imgs[tag].mean()), flush=True)
        objs, tex, sky, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{scene}.json", W, H)
        cnt = orc.ref_gpu_count(objs, cam, W, H, 256, td)
        np.savez_compressed(os.path.join(OUT, f"refpt_{scene}.npz"), W=W, H=H, spp=spp, A=imgs["A"].astype(np.float16), B=imgs["B"].astype(np.float16),
                            rays_per_sample=cnt["rays_per_sample"])
        print(scene, "rmse A/B", imgio.rmse(imgs["A"], imgs["B"]), cnt, flush=True)
