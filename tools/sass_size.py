#!/usr/bin/env python3
"""Code size of one kernel by source region (nvdisasm line info).  usage: sass_size.py object.o kernel-substring"""
import collections, os, re, subprocess, sys, tempfile
obj, key = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
def marks(path, pats):
    out = []
    for i, l in enumerate(open(path), 1):
        for k, pat in pats:
            if pat in l and k not in [o[0] for o in out]:
                out.append((k, i))
    return sorted(out, key=lambda x: x[1])
D = marks(ROOT + "/pathtracercuda_b200/csrc/trace_device.cuh", [("math", "// small math"), ("philox", "PTB_DEV uint4 philox4x32_10"), ("ld", "struct SceneView"), ("toLocal", "PTB_DEV void toLocal"),
    ("intersect", "PTB_DEV bool intersectFlat"), ("nodetest", "struct TravRay"), ("testPrim", "PTB_DEV void testPrim"), ("closestHit", "PTB_DEV Hit closestHit("), ("surface", "struct Surface"), ("tex", "PTB_DEV float4 texel"),
    ("material", "PTB_DEV V3 sampleVNDF"), ("camera", "PTB_DEV V3 cameraDir")])
cur_fn = None; cur = None; cnt = collections.Counter(); tot = 0
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if cur_fn and key in cur_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        tot += 1
        f, ln = cur if cur else ("?", 0)
        if f == "trace_device.cuh":
            name = "dev:?"
            for k, s in D:
                if ln >= s: name = "dev:" + k
        elif f.endswith(".cu"):
            name = f"{f}:{ln // 20 * 20:4d}+"
        else:
            name = f
        cnt[name] += 1
print(f"{tot} instructions = {tot * 16} bytes")
for k, c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{c:6d} {c * 16:7d} B  {k}")
