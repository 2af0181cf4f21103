#!/usr/bin/env python3
"""Bucket an ncu source-page profile by code region.  usage: ncu_buckets.py report.ncu-rep [wp|mega]
Regions are line ranges of trace_kernels.cu / trace_device.cuh found by marker comments at run time."""
import csv, io, re, subprocess, sys, os
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def marks(path, pats):
    out = {}
    for i, l in enumerate(open(path), 1):
        for k, pat in pats.items():
            if k not in out and pat in l:
                out[k] = i
    return out
K = marks(ROOT + "/pathtracercuda_b200/csrc/trace_kernels.cu", {
    "fetch": "// ---- accumulate + fetch the next pixel", "gen": "// ---- generate (trace.cu:187-192) ----", "trav": "// ---- traverse + intersect (trace.cu:112) ----",
    "miss": "// ---- environment miss (trace.cu:115-134) ----", "shade": "// ---- shade / sample (trace.cu:136-151) ----", "tail": "// ---- counters: one atomic per warp ----"})
D = marks(ROOT + "/pathtracercuda_b200/csrc/trace_device.cuh", {
    "philox": "PTB_DEV uint4 philox4x32_10", "scene": "struct SceneView", "toLocal": "PTB_DEV void toLocal", "isect": "PTB_DEV bool intersectFlat", "trav": "struct TravRay", "testprim": "// One primitive against the ray, folded", "closest": "PTB_DEV Hit closestHit(",
    "surface": "struct Surface", "tex": "PTB_DEV float4 texel", "mat": "PTB_DEV V3 sampleVNDF", "cam": "PTB_DEV V3 cameraDir"})
def bucket(f, l):
    l = int(l)
    if f == "trace_device.cuh":
        order = [("math", 0), ("philox", D["philox"]), ("ld", D["scene"]), ("toLocal", D["toLocal"]), ("intersect", D["isect"]), ("nodetest", D["trav"]), ("testPrim", D["testprim"]), ("closestHit*", D["closest"]),
                 ("surface", D["surface"]), ("tex", D["tex"]), ("material", D["mat"]), ("camera", D["cam"])]
        name = "math"
        for n, s in order:
            if l >= s: name = n
        return name
    if f == "trace_kernels.cu":
        name = "k_setup"
        for n in ["fetch", "gen", "trav", "miss", "shade", "tail"]:
            if n in K and l >= K[n]: name = "k_" + n
        return name
    return f
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; b = {}; T = [0, 0, 0]
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] not in ("", "Line No") and r[2] == "-":
        try: inst, thr, samp = int(r[7]), int(r[8]), int(r[6])
        except ValueError: continue
        x = b.setdefault(bucket(cur, r[0]), [0, 0, 0]); x[0] += inst; x[1] += thr; x[2] += samp
        T[0] += inst; T[1] += thr; T[2] += samp
for k, (i, t, s) in sorted(b.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:20s} warp-inst {100*i/T[0]:5.1f}%  thread-inst {100*t/T[1]:5.1f}%  samples {100*s/T[2]:5.1f}%  avg active thr {t/max(i,1):5.1f}  lost issue slots {100*(32*i-t)/(32*T[0]):5.1f}%")
print(f"total warp-inst {T[0]}  thread-inst {T[1]}  avg active threads {T[1]/T[0]:.2f}")
