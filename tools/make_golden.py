#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libref_host.so = the reference's own headers
built with g++, see oracle/ref_harness/ref_host.cpp).  Run in the build container, where /root/reference exists;
the fixtures are committed so boxes without the reference (the GPU box) can still pin the oracle.

  python tools/make_golden.py            # CPU fixtures (reference host build)
GPU-side fixtures (reference trace.cu on a B200) come from tools/make_golden_gpu.py run under gpurun."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
from oracle import imgio, orc

G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)
R = orc.RefHost()
rng = np.random.default_rng(20260101)


def objs_to_array(objs):
    return np.frombuffer(bytes(pt.object_array(objs))[: len(objs) * 80], np.uint8).copy()


def scene_fixture(name, objs, cam, W, H):
    R.set_scene(objs)
    R.set_camera(cam)
    idx, t, st = R.primary_pass(W, H)
    nodes, depth, valid = R.bvh_info()
    rows = np.zeros((len(objs), 12), np.float32)
    boxes = np.zeros((len(objs), 6), np.float32)
    for i in range(len(objs)):
        r, _, b, _ = R.object_bytes(i)
        rows[i], boxes[i] = r, b
    # secondary-style rays: origins on surfaces (primary hit points), random directions
    hit = np.nonzero(idx >= 0)[0]
    sel = rng.choice(hit, size=min(4000, len(hit)), replace=False)
    o = np.zeros((len(sel), 3), np.float32)
    d = rng.normal(size=(len(sel), 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    for k, p in enumerate(sel):
        x, y = p % W, p // W
        ray = R.camera_ray(np.float32((x + 0.5) / W), np.float32((y + 0.5) / H))
        o[k] = ray[:3] + t[p] * ray[3:]
    si, st_, sn = R.trace_rays(o, d)
    np.savez_compressed(os.path.join(G, f"{name}.npz"), objects=objs_to_array(objs), camera=np.frombuffer(bytes(cam), np.uint8).copy(), W=W, H=H,
                        primary_idx=idx, primary_t=t, node_visits=st[0], prim_tests=st[1], bvh_nodes=nodes, bvh_depth=depth, bvh_valid=valid,
                        rows=rows, boxes=boxes, ref_camera=R.get_camera(), sec_o=o, sec_d=d, sec_idx=si, sec_t=st_, sec_n=sn)
    print(name, len(objs), "objects", W, H, "hit fraction", float((idx >= 0).mean()), "nodes", nodes, "depth", depth)


# 1. the two bundled scenes through the reference's OWN loader (SceneLoader.cpp) vs our Python schema mirror
for scene, (W, H) in {"cornell_box": (96, 96), "generated_scene": (192, 108)}.items():
    path = f"{pt.ASSETS}/scenes/{scene}.json"
    objs, tex, sky, cam = pt.parse_scene_py(path, W, H)
    n = R.load_scene_file(path, W, H, cwd=pt.ASSETS)
    assert n == len(objs)
    for i in range(n):  # the reference loader and our schema mirror must build identical objects
        R2 = R.object_bytes(i)
    ref_rows = np.array([R.object_bytes(i)[0] for i in range(n)])
    ref_cam = R.get_camera()
    scene_fixture(scene, objs, cam, W, H)
    d = np.load(os.path.join(G, f"{scene}.npz"))
    assert np.array_equal(d["rows"].view(np.uint32), ref_rows.view(np.uint32)), "schema mirror != reference loader"
    assert np.array_equal(d["ref_camera"].view(np.uint32), ref_cam.view(np.uint32))

# 2. synthetic mixed scene: every shape, rotation, material
objs, cam = scenegen.synthetic_scene(1500, 160, 90)
scene_fixture("synthetic_1500", objs, cam, 160, 90)

# 3. per-shape hit probes (Hittable::hit) on a small rotated/scaled zoo, incl. rays starting inside shapes
zoo = [pt.make_object(s, position=(i * 0.3 - 1, 0.2 * i, -0.1 * i), rotation_deg=(20 * i, -35 * i, 50 * i), scale=(0.5 + 0.1 * i, 0.7, 0.4 + 0.05 * i))
       for i, s in enumerate(pt.SHAPES)]
R.set_scene(zoo)
N = 6000
o = (rng.random((N, 3)).astype(np.float32) * 4 - 2)
o[::5] *= 0.15  # many origins inside / near the shapes
tgt = (rng.random((N, 3)).astype(np.float32) * 1.6 - 0.8)
d = tgt - o
d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
which = rng.integers(0, 7, N)
res = np.full((N, 10), np.nan, np.float32)
hitf = np.zeros(N, np.int32)
tmax = np.where(rng.random(N) < 0.3, rng.random(N) * 3, 3.4028234663852886e38).astype(np.float32)
for k in range(N):
    r = R.hit_object(int(which[k]), o[k], d[k], 0.001, float(tmax[k]))
    if r is not None:
        res[k] = r
        hitf[k] = 1
np.savez_compressed(os.path.join(G, "hit_probes.npz"), objects=objs_to_array(zoo), o=o, d=d, which=which, tmax=tmax, hit=hitf, res=res)
print("hit probes", N, "hits", int(hitf.sum()))

# 4. Material::sample vectors at the reference's own XORWOW draws
M = 4000
mats = np.zeros((M, 8), np.float32)  # type, base rgb, roughness, metalness, seed, pad
Ns = np.zeros((M, 3), np.float32)
ins = np.zeros((M, 3), np.float32)
outs = np.zeros((M, 13), np.float32)
for k in range(M):
    mt = int(rng.integers(0, 3))
    bc = rng.random(3)
    rough = float(rng.random()) if k % 7 else 0.01
    metal = float(rng.integers(0, 2)) if rng.random() < 0.7 else float(rng.random())
    m = pt.make_object("SPHERE", material=mt, base_color=bc, roughness=rough, metalness=metal, emissive=(0.5, 0.25, 2.0) if k % 11 == 0 else (0, 0, 0)).material
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)
    if k % 50 == 0:
        n = np.array([0, 0, 1.0]) * (1 if k % 100 else -1)
    dd = rng.normal(size=3)
    dd /= np.linalg.norm(dd)
    if np.dot(dd, n) > 0:
        dd = -dd
    seed = int(rng.integers(0, 2 ** 31))
    Ns[k], ins[k] = n, dd
    outs[k] = R.material_sample(m, Ns[k], ins[k], seed)
    mats[k] = [mt, *m.base_color, m.roughness, m.metalness, 0, 1 if k % 11 == 0 else 0]
np.savez_compressed(os.path.join(G, "material_samples.npz"), mats=mats, N=Ns, in_dir=ins, out=outs)
print("material samples", M, "nan rows", int(np.isnan(outs).any(1).sum()))

# 5. full path tracing by the reference host build (XORWOW seeds 1984+pixel and 7919+pixel): the noise-floor pair
sky = imgio.read_hdr(pt.ASSETS + "/skybox.hdr")
for scene, (W, H, spp) in {"cornell_box": (48, 48, 2048), "generated_scene": (64, 36, 2048)}.items():
    path = f"{pt.ASSETS}/scenes/{scene}.json"
    objs, tex, skyi, cam = pt.parse_scene_py(path, W, H)
    for ob in objs:
        ob.material.texture = 0  # the host build cannot sample base-colour textures (reference compiles that out, Material.inl:26)
    R.set_scene(objs)
    R.set_camera(cam)
    R.set_sky(sky if scene == "generated_scene" else None)
    a, ra = R.render(W, H, spp, 1984)
    b, rb = R.render(W, H, spp, 7919)
    np.savez_compressed(os.path.join(G, f"render_host_{scene}.npz"), W=W, H=H, spp=spp, seedA=(a[..., :3] / spp).astype(np.float32),
                        seedB=(b[..., :3] / spp).astype(np.float32), raysA=ra, raysB=rb)
    print("render", scene, W, H, spp, "rays/sample", ra / (W * H * spp), "rmse A/B", imgio.rmse(a / spp, b / spp))

# 6. tonemap vectors (kernels/tonemap.cu arithmetic on the host)
acc = (rng.random((4096, 4)).astype(np.float32) ** 3) * 4000
acc[:16] = 0
np.savez_compressed(os.path.join(G, "tonemap.npz"), accum=acc, count=np.int32(37), out=R.tonemap(acc, 37))
print("done")
