#!/usr/bin/env python3
"""BASELINE config 4: synthetic procedural scenes of N mixed quadrics at 1080p (BVH exceeds shared memory: L2/HBM-bound
traversal).  Times our trace kernel and the reference program (oracle/_ref/ref_pt) on the same scene file.
usage: tools/large_scenes.py [N,N,...] [spp]"""
import json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
from oracle import orc
Ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10000, 100000, 1000000]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = 1920, 1080
for n in Ns:
    with tempfile.TemporaryDirectory() as td:
        os.symlink(pt.ASSETS + "/skybox.hdr", td + "/skybox.hdr")
        f = td + "/scene.json"
        t0 = time.perf_counter(); scenegen.write_synthetic_scene(f, n); tgen = time.perf_counter() - t0
        out = {"objects": n + 1, "spp": spp}
        with pt.Pathtracer(W, H) as P:
            t0 = time.perf_counter()
            cam = P.loadSceneFile(f, cwd=td)
            out["load_parse_build_upload_s"] = round(time.perf_counter() - t0, 3)
            P.render(cam, 4, True)
            best = min((P.render(cam, spp, True), P.getTiming())[1] for _ in range(2))
            st = P.stats()
            out.update({"ours_ms": round(best, 2), "ours_Mrays_s": round(st.rays / best / 1e3, 1), "ours_Msamples_s": round(st.samples / best / 1e3, 1), "rays_per_sample": round(st.rays / st.samples, 3),
                        "bvh_nodes": st.bvh_nodes, "bvh_depth": st.bvh_depth, "scene_MB": round(st.scene_bytes / 1e6, 2), "scene_in_smem": st.scene_in_smem})
            P.setOption("count_work", 1); P.render(cam, 4, True); s2 = P.stats(); P.setOption("count_work", 0)
            out.update({"nodes_per_ray": round(s2.node_visits / s2.rays, 2), "prims_per_ray": round(s2.prim_tests / s2.rays, 2)})
        if os.path.exists(orc.REF_PT) and "--no-ref" not in sys.argv:
            p = subprocess.run([orc.REF_PT, "-w", str(W), "-h", str(H), "-spp", str(spp), "scene.json"], cwd=td, capture_output=True, text=True)
            ms = [float(l.split(" in ")[1].split(" ms")[0]) for l in p.stdout.splitlines() if l.startswith("Finished accumulating")]
            if ms:
                out.update({"ref_ms": round(ms[0], 2), "ref_Msamples_s": round(W * H * spp / ms[0] / 1e3, 1), "speedup_samples": round(ms[0] / best, 2)})
            else:
                out["ref_error"] = (p.stderr or p.stdout)[-200:]
        print(json.dumps(out), flush=True)
