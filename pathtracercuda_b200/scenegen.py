"""Synthetic procedural scenes (BASELINE config 4): N mixed quadrics on a jittered grid over a floor quad.

Follows the recipe of the reference's own (compiled-out) generator, generateSceneFile (reference
SceneLoader.cpp:12-122): floor QUAD scaled 20x, objects on an integer grid with 0.9 jitter at height 0.2, albedo =
product of two uniform RGBs, metalness in {0,1} with p = 1/2, roughness U(0,1) - widened as SURVEY.md §8d asks:
uniform over all 7 shapes AND all 3 material types, random Euler rotations U(-180,180) degrees, scale U(0.1,0.3),
grid side ceil(sqrt(N)), seed 1984.  Deterministic (numpy PCG64)."""
import json
import math

import numpy as np

from .abi import MATERIALS, SHAPES, make_camera, make_object

SHAPE_NAMES = list(SHAPES.keys())
MATERIAL_NAMES = list(MATERIALS.keys())


def synthetic_scene_dict(n, seed=1984, skybox="skybox.hdr"):
    rng = np.random.default_rng(seed)
    side = int(math.ceil(math.sqrt(n)))
    half = side / 2.0
    objs = [{"name": "floor", "type": "QUAD", "position": [0.0, 0.0, 0.0], "rotation": [0.0, 0.0, 0.0],
             "scale": [float(half + 2.0), 1.0, float(half + 2.0)],
             "material": {"type": "LAMBERT", "baseColor": [1.0, 1.0, 1.0], "emissive": [0.0, 0.0, 0.0], "roughness": 1.0, "metalness": 0.0, "texture": ""}}]
    u = rng.random((n, 16))
    for i in range(n):
        a, b = i % side, i // side
        r = u[i]
        albedo = [float(r[2] * r[5]), float(r[3] * r[6]), float(r[4] * r[7])]
        objs.append({
            "name": "", "type": SHAPE_NAMES[int(r[8] * 7) % 7],
            "position": [float(a - half + 0.9 * r[0]), 0.2 + 0.1, float(b - half + 0.9 * r[1])],
            "rotation": [float(r[9] * 360 - 180), float(r[10] * 360 - 180), float(r[11] * 360 - 180)],
            "scale": [float(0.1 + 0.2 * r[12])] * 3,
            "material": {"type": MATERIAL_NAMES[int(r[13] * 3) % 3], "baseColor": albedo, "emissive": [0.0, 0.0, 0.0],
                         "roughness": float(r[14]), "metalness": 1.0 if r[15] > 0.5 else 0.0, "texture": ""}})
    d = float(max(half, 4.0))
    return {"camera": {"position": [d * 1.3, d * 0.45 + 1.0, d * 0.3], "look_at": [0.0, 0.0, 0.0], "fovy": 60.0}, "skybox": skybox, "objects": objs}


def write_synthetic_scene(path, n, seed=1984, skybox="skybox.hdr"):
    """the scene of synthetic_scene_dict as a JSON file - byte for byte what json.dump(..., separators=(",", ":")) writes, put
    together by hand (a million-object scene through json.dump takes a minute)"""
    rng = np.random.default_rng(seed)
    side = int(math.ceil(math.sqrt(n)))
    half = side / 2.0
    d = float(max(half, 4.0))
    r_ = repr
    head = ('{"camera":{"position":[%s,%s,%s],"look_at":[0.0,0.0,0.0],"fovy":60.0},"skybox":%s,"objects":[' % (r_(d * 1.3), r_(d * 0.45 + 1.0), r_(d * 0.3), json.dumps(skybox))
            + '{"name":"floor","type":"QUAD","position":[0.0,0.0,0.0],"rotation":[0.0,0.0,0.0],"scale":[%s,1.0,%s],' % (r_(float(half + 2.0)), r_(float(half + 2.0)))
            + '"material":{"type":"LAMBERT","baseColor":[1.0,1.0,1.0],"emissive":[0.0,0.0,0.0],"roughness":1.0,"metalness":0.0,"texture":""}}')
    u = rng.random((n, 16))
    idx = np.arange(n)
    px = (idx % side) - half + 0.9 * u[:, 0]
    pz = (idx // side) - half + 0.9 * u[:, 1]
    alb = u[:, 2:5] * u[:, 5:8]
    rot = u[:, 9:12] * 360 - 180
    sc = 0.1 + 0.2 * u[:, 12]
    shape = (u[:, 8] * 7).astype(np.int64) % 7
    mat = (u[:, 13] * 3).astype(np.int64) % 3
    with open(path, "w") as f:
        f.write(head)
        chunk = []
        for i in range(n):
            s = r_(float(sc[i]))
            chunk.append(',{"name":"","type":"%s","position":[%s,0.30000000000000004,%s],"rotation":[%s,%s,%s],"scale":[%s,%s,%s],'
                         '"material":{"type":"%s","baseColor":[%s,%s,%s],"emissive":[0.0,0.0,0.0],"roughness":%s,"metalness":%s,"texture":""}}'
                         % (SHAPE_NAMES[shape[i]], r_(float(px[i])), r_(float(pz[i])), r_(float(rot[i, 0])), r_(float(rot[i, 1])), r_(float(rot[i, 2])), s, s, s,
                            MATERIAL_NAMES[mat[i]], r_(float(alb[i, 0])), r_(float(alb[i, 1])), r_(float(alb[i, 2])), r_(float(u[i, 14])), "1.0" if u[i, 15] > 0.5 else "0.0"))
            if len(chunk) == 65536:
                f.write("".join(chunk))
                chunk = []
        f.write("".join(chunk))
        f.write("]}\n")


def synthetic_scene(n, width, height, seed=1984):
    """-> (objects, camera) as ABI structs, without going through a file"""
    d = synthetic_scene_dict(n, seed)
    objs = []
    for o in d["objects"]:
        m = o["material"]
        objs.append(make_object(o["type"], o["position"], o["rotation"], o["scale"], m["type"], m["baseColor"], m["emissive"], m["roughness"], m["metalness"], 0))
    c = d["camera"]
    return objs, make_camera(c["position"], c["look_at"], c["fovy"], np.float32(width) / np.float32(height))
