"""Host-side logic of the product library that needs no GPU: the JSON scene loader, the image codecs, the scene
compiler's BVH, the C-ABI surface and the CLI's argument handling."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import pathtracercuda_b200 as pt
from pathtracercuda_b200 import api, scenegen
from oracle import imgio, orc
from tests.emu.emu import Emu
from tests.helpers import ROOT, bits, golden_objects, load_golden

CLI = os.path.join(ROOT, "pathtracercuda_b200", "bin", "pathtracer_b200")


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pt_b200.h")).read()
    declared = set(re.findall(r"\b(pt_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    L = pt.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.pt_version()
    assert C.sizeof(pt.ObjectDesc) == 80 and C.sizeof(pt.CameraDesc) == 44 and C.sizeof(pt.MaterialDesc) == 40


@pytest.mark.skipif(__import__("tests.conftest", fromlist=["HAS_GPU"]).HAS_GPU, reason="box has a GPU")
def test_no_cpu_fallback():
    with pytest.raises(pt.PtError, match="no CUDA device"):
        pt.Pathtracer(16, 16)


def test_camera_controls_match_reference():
    """pt_camera_rotate / pt_camera_translate against the reference's own Camera::rotate / translate (Camera.inl:30-52,
    fixture from tools/make_golden_camera.py): after every move of a 24-move walk the camera rebuilt from our description
    has the reference's origin, image-plane vectors and basis (the description round-trips the basis, so a few ulp per move)"""
    d = load_golden("camera_moves")
    for k in range(d["start"].shape[0]):
        from pathtracercuda_b200.abi import CameraDesc
        cam = CameraDesc.from_buffer_copy(bytes(d["start"][k]))
        O = orc.Oracle([pt.make_object("SPHERE")])
        for m, ref in zip(d["moves"][k], d["states"][k][1:]):
            if m[0] == 0:
                pt.camera_rotate(cam, float(m[1]), float(m[2]), float(m[3]))
            else:
                pt.camera_translate(cam, float(m[1]), float(m[2]), float(m[3]))
            origin, ll, hor, ver = ref[2:5], ref[5:8], ref[8:11], ref[11:14]
            assert np.allclose(np.array(cam.position), origin, rtol=0, atol=2e-5)
            for s_, t_ in [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (0.5, 0.5), (0.9, 0.2)]:
                ray = O.camera_ray(cam, s_, t_)  # Camera ctor + getRay on OUR description
                v = ll + np.float32(s_) * hor + np.float32(t_) * ver
                assert np.allclose(ray[3:], v / np.linalg.norm(v), rtol=0, atol=5e-6), (k, m)
        assert not np.allclose(np.array(cam.position), d["states"][k][0][2:5])  # the walk went somewhere


def test_large_scene_file_is_parsed_by_all_cores(tmp_path):
    """an `objects` list of several megabytes is split at its top-level commas and parsed / converted by all cores
    (json_min.cpp parallelArray, scene_loader.cpp): same objects, texture handles in first-use order, messages in object
    order, the same error texts as the one-thread path"""
    n = 20000
    f = tmp_path / "big.json"
    scenegen.write_synthetic_scene(str(f), n)
    d = json.loads(f.read_text())
    d["objects"][777]["material"]["texture"] = "b.png"
    d["objects"][123]["material"]["texture"] = "a.png"
    d["objects"][9000]["material"]["texture"] = "b.png"
    d["objects"][15000]["type"] = "TORUS, [not] a {shape}"
    d["objects"][15001]["name"] = 'brackets ] } , " in a string \\'
    f.write_text(json.dumps(d, indent=1))
    assert f.stat().st_size > 5 << 20
    objs, paths, sky, cam = pt.parse_scene_file(str(f), 64, 36)
    o2, p2, s2, c2 = pt.parse_scene_py(str(f), 64, 36)
    assert len(objs) == n + 1 and paths == p2 and paths[:2] == ["a.png", "b.png"] and sky == s2 and bytes(cam) == bytes(c2)
    assert all(bytes(a) == bytes(b) for a, b in zip(objs, o2))
    assert objs[123].material.texture == 1 and objs[777].material.texture == 2 and objs[9000].material.texture == 2
    # one thread (the sequential path) gives the same bytes
    code = ("import sys; sys.path.insert(0, %r); import pathtracercuda_b200 as pt; o, p, s, c = pt.parse_scene_file(%r, 64, 36); "
            "import hashlib; print(hashlib.sha1(b''.join(bytes(x) for x in o)).hexdigest(), p, s)" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(f)))
    one = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
    import hashlib
    assert one.returncode == 0, one.stderr
    assert one.stdout.split()[0] == hashlib.sha1(b"".join(bytes(x) for x in objs)).hexdigest()
    # errors inside the big list: a syntax error is reported with its byte offset by the sequential path, a type error by the loader
    text = f.read_text()
    k = text.index('"position"', len(text) // 2)
    f.write_text(text[:k] + '"position": [1, 2, oops],' + text[k:])
    with pytest.raises(pt.PtError, match=r"JSON parse error at byte \d+: unexpected character"):
        pt.parse_scene_file(str(f), 4, 4)
    d["objects"][12000]["scale"] = [1, "two", 3]
    d["objects"][18000]["position"] = ["x", 0, 0]
    f.write_text(json.dumps(d))
    with pytest.raises(pt.PtError, match='type must be number in "scale"'):
        pt.parse_scene_file(str(f), 4, 4)


@pytest.mark.parametrize("scene", ["cornell_box", "generated_scene"])
def test_scene_loader_matches_schema_mirror_and_reference(scene):
    path = f"{pt.ASSETS}/scenes/{scene}.json"
    objs, paths, sky, cam = pt.parse_scene_file(path, 1920, 1080)
    o2, p2, s2, c2 = pt.parse_scene_py(path, 1920, 1080)
    assert paths == p2 and sky == s2 and bytes(cam) == bytes(c2)
    assert len(objs) == len(o2) and all(bytes(a) == bytes(b) for a, b in zip(objs, o2))
    # and against the objects the REFERENCE's own loader produced (golden fixture, via libref_host)
    g, gcam = golden_objects(load_golden(scene))
    assert all(bytes(a) == bytes(b) for a, b in zip(objs, g))


def test_scene_loader_quirks(tmp_path):
    scene = {"camera": {"position": [1, 2, 3], "fovy": 40}, "skybox": "",
             "objects": [{"type": "TORUS", "rotation": [90, 0, 0], "material": {"type": "GLASS", "roughness": 1, "metalness": 0.25, "texture": "a.png"}},
                         {"type": "DISK", "scale": [2.0, 3.0, 4.0], "material": {"texture": "a.png", "emissive": [1, 2, 3]}}, {}]}
    p = tmp_path / "s.json"
    p.write_text(json.dumps(scene))
    objs, paths, sky, cam = pt.parse_scene_file(str(p), 200, 100)
    assert len(objs) == 3 and paths == ["a.png"] and sky == 0
    assert objs[0].type == pt.SHAPES["SPHERE"] and objs[0].material.type == pt.MATERIALS["LAMBERT"]  # unknown names keep defaults
    assert objs[0].material.roughness == 0.5  # integer literal ignored (Q5) ...
    assert objs[0].material.metalness == 0.25  # ... float literal read
    assert abs(objs[0].rotation[0] - np.float32(np.pi / 2)) < 1e-6
    assert objs[0].material.texture == 1 and objs[1].material.texture == 1  # de-duplicated by path
    assert list(objs[1].material.emissive) == [1.0, 2.0, 3.0] and list(objs[1].scale) == [2.0, 3.0, 4.0]
    assert objs[2].type == 0 and list(objs[2].scale) == [1.0, 1.0, 1.0]
    assert list(cam.position) == [1.0, 2.0, 3.0] and list(cam.look_at) == [0.0, 0.0, -1.0]
    assert abs(cam.fovy - np.float32(np.pi / 3)) < 1e-6 and cam.aspect == 2.0  # "fovy": 40 is an int literal -> default 60
    for bad in ["{", '{"objects": [1,]}', '{"a": 01}', "[1] x"]:
        p.write_text(bad)
        with pytest.raises(pt.PtError, match="JSON parse error"):
            pt.parse_scene_file(str(p), 4, 4)
    with pytest.raises(pt.PtError, match="Failed to open input file"):
        pt.parse_scene_file(str(tmp_path / "missing.json"), 4, 4)
    # the file is mapped, not read: an empty file, blanks only, a directory
    for content in ("", "   \n"):
        p.write_text(content)
        with pytest.raises(pt.PtError, match="JSON parse error at byte \\d+: unexpected end of input"):
            pt.parse_scene_file(str(p), 4, 4)
    with pytest.raises(pt.PtError, match="JSON parse error"):
        pt.parse_scene_file(str(tmp_path), 4, 4)
    p.write_text('{"objects": [{"position": ["x", 0, 0]}]}')
    with pytest.raises(pt.PtError, match="type must be number"):
        pt.parse_scene_file(str(p), 4, 4)
    p.write_text('﻿{"skybox": "sky\\u0041.hdr", "objects": []}')
    objs, paths, sky, cam = pt.parse_scene_file(str(p), 4, 4)
    assert objs == [] and paths == ["skyA.hdr"] and sky == 1


def test_image_codecs(tmp_path):
    # decoders vs independent ones (PIL / numpy)
    assert np.array_equal(pt.read_image(pt.ASSETS + "/earth.png"), imgio.read_png(pt.ASSETS + "/earth.png"))
    assert np.array_equal(bits(pt.read_image(pt.ASSETS + "/skybox.hdr")), bits(imgio.read_hdr(pt.ASSETS + "/skybox.hdr")))
    rng = np.random.default_rng(3)
    from PIL import Image
    for mode, arr in {"L": rng.integers(0, 256, (7, 5), np.uint8), "LA": rng.integers(0, 256, (7, 5, 2), np.uint8),
                      "RGB": rng.integers(0, 256, (9, 13, 3), np.uint8), "RGBA": rng.integers(0, 256, (9, 13, 4), np.uint8)}.items():
        f = str(tmp_path / f"{mode}.png")
        Image.fromarray(arr, mode).save(f)
        assert np.array_equal(pt.read_image(f), imgio.read_png(f)), mode
    pal = Image.fromarray(rng.integers(0, 256, (8, 8, 3), np.uint8), "RGB").quantize(16)
    pal.save(str(tmp_path / "pal.png"))
    assert np.array_equal(pt.read_image(str(tmp_path / "pal.png")), imgio.read_png(str(tmp_path / "pal.png")))
    (tmp_path / "bad.png").write_bytes(b"not a png")
    with pytest.raises(pt.PtError):
        pt.read_image(str(tmp_path / "bad.png"))
    # writers: flip on write (buffer row 0 = bottom of the view), PNG lossless, HDR = stb's truncating RGBE
    img8 = rng.integers(0, 256, (11, 17, 4), np.uint8)
    img8[..., 3] = 255
    f = str(tmp_path / "o.png")
    pt.write_png(f, img8)
    assert np.array_equal(imgio.read_png(f), img8[::-1])
    for w in (5, 40):  # flat (w < 8) and run-length encoded scanlines
        imgf = (rng.random((6, w, 4)).astype(np.float32) ** 4) * 50
        imgf[1, :, :] = 0.25  # long runs
        imgf[2, 0] = 0
        f = str(tmp_path / f"o{w}.hdr")
        pt.write_hdr(f, imgf)
        back = imgio.read_hdr(f)[::-1]
        m = imgf[..., :3].max(-1)
        e = np.frexp(m)[1]
        expect = np.floor(imgf[..., :3] * (np.frexp(m)[0] * 256.0 / np.maximum(m, 1e-38))[..., None]) * np.ldexp(1.0, e - 8)[..., None]
        expect[m < 1e-32] = 0
        assert np.allclose(back[..., :3], expect, rtol=1e-6, atol=0)
        assert np.array_equal(bits(pt.read_image(f)), bits(imgio.read_hdr(f)))  # our reader on our writer
        assert open(f, "rb").read().startswith(b"#?RADIANCE\n# Written by stb_image_write.h\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=          1.0000000000000\n\n-Y 6 +X ")


def _area(b):
    e = np.maximum(b[1] - b[0], 0)
    return 2 * (e[0] * e[1] + e[0] * e[2] + e[1] * e[2])


def _validate_bvh(objs, max_leaf=4, builder=0):
    E = Emu(objs, max_leaf, builder)
    nodes, prims = E.arrays()
    O = orc.Oracle(objs)
    boxes = np.array([O.object_info(i)[1] for i in range(len(objs))])
    nf = nodes.view(np.float32)
    seen = np.zeros(len(objs), int)
    # the fourth quad of a primitive record (pt_types.h): quadric coefficients B, H, J as floats, then one packed word
    packed = prims[:, 15].astype(np.uint32)
    scene_of, type_of = (packed & 0xffffff).astype(np.int64), ((packed >> 24) & 7).astype(np.int64)
    assert sorted(scene_of.tolist()) == list(range(len(objs)))  # sceneIndex is a permutation
    coef = prims[:, 12:15].copy().view(np.float32)
    for k in range(len(objs)):
        t = int(objs[int(scene_of[k])].type)
        assert type_of[k] == t
        want = {0: (1, 0, -1), 1: (0, 0, -1), 3: (-1, 0, 0), 4: (0, -1, 0)}.get(t, (0, 0, 0))  # sphere, cylinder, cone, paraboloid (Hittable.inl:151,176,242,273)
        assert tuple(coef[k]) == want
        bits = int(packed[k]) >> 27
        assert bits == (int(objs[int(scene_of[k])].material.texture != 0) | (2 if t in (2, 5) else 0) | (4 if t == 6 else 0) | (8 if t == 2 else 0) | (16 if t == 0 else 0))

    def walk(ref, depth):
        if ref < 0:
            u = ref & 0xffffffff
            first, count, typ = u & 0xffffff, (u >> 24) & 15, (u >> 28) & 7
            assert 1 <= count <= max(max_leaf, 1)
            assert typ == type_of[first]
            cls = [1 if type_of[k] in (2, 5) else (2 if type_of[k] == 6 else 0) for k in range(first, first + count)]
            assert cls == sorted(cls), "leaf primitives are ordered quadrics, flat shapes, cubes"
            b = np.array([[np.inf] * 3, [-np.inf] * 3])
            for k in range(first, first + count):
                s = scene_of[k]
                seen[s] += 1
                b[0] = np.minimum(b[0], boxes[s][:3])
                b[1] = np.maximum(b[1], boxes[s][3:])
            return b, depth
        n = nf[ref]
        kids = nodes[ref, 12:14].view(np.int32)
        out = np.array([[np.inf] * 3, [-np.inf] * 3])
        dmax = depth
        for c in range(2):
            if kids[c] == -2 ** 31:
                continue
            b, d = walk(int(kids[c]), depth + 1)
            ctr, half = n[[c, 2 + c, 4 + c]].astype(np.float64), n[[6 + c, 8 + c, 10 + c]].astype(np.float64)  # pt_types.h: children interleaved
            stored = np.array([ctr - half, ctr + half])      # nodes hold centre / (padded) half extent
            assert (half < (b[1] - b[0]) * 0.5 * (1 + 1e-5) + 1e-4).all(), "padding must stay tiny"
            assert (stored[0] <= b[0]).all() and (stored[1] >= b[1]).all(), "child box must contain its primitives"
            out[0] = np.minimum(out[0], b[0])                # exact bounds of the subtree (the padding is per box, not nested)
            out[1] = np.maximum(out[1], b[1])
            dmax = max(dmax, d)
        return out, dmax

    _, depth = walk(0, 0)
    # primitives that span the scene are hoisted out of the tree and tested by every ray first: the first G records
    G = E.global_count()
    assert G <= 8 and G < len(objs)
    scene_area = _area(np.array([boxes[:, :3].min(0), boxes[:, 3:].max(0)]))
    for k in range(len(objs)):
        a = _area(np.array([boxes[scene_of[k]][:3], boxes[scene_of[k]][3:]]))
        if k < G:
            seen[scene_of[k]] += 1
            assert a >= 0.25 * scene_area * (1 - 1e-5)
        elif G < 8 and G < len(objs) - 1:
            assert a < 0.25 * scene_area * (1 + 1e-5), "a scene-spanning primitive was left in the tree"
    assert (seen == 1).all(), "every primitive in exactly one leaf or hoisted (reference BVH::validate)"
    info = E.info()
    assert info[0] == len(nodes) and depth <= info[1] + 1
    return info


def test_bvh_structure():
    for name in ("cornell_box", "generated_scene"):
        objs, _ = golden_objects(load_golden(name))
        _validate_bvh(objs)
    objs, _ = scenegen.synthetic_scene(3000, 64, 36)
    nodes, depth, leaves = _validate_bvh(objs)
    assert depth < 40
    _validate_bvh([pt.make_object("SPHERE")])                        # a single object: root with one empty child
    _validate_bvh([pt.make_object("CUBE")] * 9, max_leaf=2)          # coincident centroids: median fallback
    _validate_bvh([pt.make_object("QUAD", position=(i, 0, 0)) for i in range(5)], max_leaf=8)
    # the LBVH builder (Morton order + Karras' radix tree, option "bvh_builder" = 1): the same invariants, one primitive per leaf,
    # also with coincident centroids (equal Morton keys: ties broken by position) and the smallest trees
    for name in ("cornell_box", "generated_scene"):
        objs, _ = golden_objects(load_golden(name))
        _validate_bvh(objs, max_leaf=1, builder=1)
    objs, _ = scenegen.synthetic_scene(3000, 64, 36)
    n2, d2, l2 = _validate_bvh(objs, max_leaf=1, builder=1)
    assert d2 < 46 and l2 >= leaves
    _validate_bvh([pt.make_object("SPHERE")], builder=1)
    _validate_bvh([pt.make_object("SPHERE"), pt.make_object("CUBE", position=(3, 0, 0))], max_leaf=1, builder=1)
    _validate_bvh([pt.make_object("CUBE")] * 9, max_leaf=1, builder=1)
    _validate_bvh([pt.make_object("QUAD", position=(i, 0, 0)) for i in range(5)], max_leaf=1, builder=1)
    # centroids that crowd towards a point along every axis (one-hot Morton keys): the radix tree is a chain deeper than the
    # traversal stack - it is thrown away and the SAH builder takes over
    crowd = [pt.make_object("SPHERE", position=tuple(2.0 ** -k if a == ax else 0.0 for a in range(3)), scale=(1e-9,) * 3) for k in range(21) for ax in range(3)]
    crowd.append(pt.make_object("SPHERE", scale=(1e-9,) * 3))
    assert _validate_bvh(crowd, max_leaf=1, builder=1) == _validate_bvh(crowd, max_leaf=1, builder=0)


def test_cli_arguments():
    assert os.path.exists(CLI), "build the CLI with make -C pathtracercuda_b200/csrc"
    # like the reference, options are only parsed in argv[1..argc-2]; the last argument is always the scene path
    # (main.cpp:40,146-149), so a lone "-help" is taken as the input file and help needs a second argument
    r = subprocess.run([CLI, "-help", "x.json"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("USAGE: PathtracerCUDA.exe [options] <input file>")
    for opt in ("-help", "-w", "-h", "-spp", "-window", "-enable_controls", "-o", "-ohdr"):
        assert re.search(rf"^{opt}\s", r.stdout, re.M), opt
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 0 and "Missing input file argument!" in r.stdout
    r = subprocess.run([CLI, "-w", "0", "x.json"], capture_output=True, text=True)
    assert r.returncode == 0 and "Invalid input for -w!" in r.stdout
    r = subprocess.run([CLI, "-bogus", "x.json"], capture_output=True, text=True)
    assert r.returncode == 0 and "Can't parse argument: -bogus" in r.stdout
    # the double-dash extras are understood (what happens next needs a GPU: here the context cannot be created)
    r = subprocess.run([CLI, "--env-is", "--bvh", "lbvh", "--seed", "7", "--stats", "-w", "8", "-h", "8", "x.json"], capture_output=True, text=True)
    assert "Can't parse argument" not in r.stdout and "Width: 8" in r.stdout


def test_synthetic_generator_deterministic(tmp_path):
    a = scenegen.synthetic_scene_dict(200)
    b = scenegen.synthetic_scene_dict(200)
    assert a == b and len(a["objects"]) == 201
    kinds = {o["type"] for o in a["objects"]}
    mats = {o["material"]["type"] for o in a["objects"]}
    assert kinds == set(pt.SHAPES) and mats == set(pt.MATERIALS)
    f = tmp_path / "syn.json"
    scenegen.write_synthetic_scene(str(f), 200)
    objs, paths, sky, cam = pt.parse_scene_file(str(f), 64, 36)
    o2, c2 = scenegen.synthetic_scene(200, 64, 36)
    assert all(bytes(x) == bytes(y) for x, y in zip(objs, o2))
