"""The device functions of the CUDA path (pathtracercuda_b200/csrc/trace_device.cuh), compiled for the host by the
TEST-ONLY emulation build (tests/emu), against the oracle.  This is how the kernel LOGIC is checked on a box without a
GPU; the real kernels are checked by the `-m gpu` tests."""
import numpy as np
import pytest

import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
from oracle import imgio, orc
from tests.emu.emu import Emu
from tests.helpers import bits, golden_objects, load_golden


@pytest.mark.parametrize("name", ["cornell_box", "generated_scene", "synthetic_1500"])
@pytest.mark.parametrize("max_leaf", [1, 4])
def test_primary_and_secondary_parity(name, max_leaf):
    d = load_golden(name)
    objs, cam = golden_objects(d)
    E = Emu(objs, max_leaf)
    O = orc.Oracle(objs)
    W, H = int(d["W"]), int(d["H"])
    idx, t, _ = E.primary_pass(cam, W, H)
    hit = (idx >= 0) & (d["primary_idx"] >= 0)
    assert np.array_equal(idx >= 0, d["primary_idx"] >= 0)
    assert (np.abs(t - d["primary_t"])[hit] <= 1e-5 * d["primary_t"][hit]).all()
    for p in np.nonzero(idx != d["primary_idx"])[0]:  # only exact geometric ties may differ
        x, y = p % W, p // W
        ray = O.camera_ray(cam, np.float32((x + 0.5) / W), np.float32((y + 0.5) / H))
        a, b = O.hit_object(int(idx[p]), ray[:3], ray[3:]), O.hit_object(int(d["primary_idx"][p]), ray[:3], ray[3:])
        assert a is not None and b is not None and abs(a[0] - b[0]) <= 1e-5 * a[0]
    si, st, sn = E.trace_rays(d["sec_o"], d["sec_d"])
    same = si == d["sec_idx"]
    assert same.mean() > 0.998
    h = same & (si >= 0)
    assert (np.abs(st - d["sec_t"])[h] <= 1e-5 * np.maximum(d["sec_t"][h], 1e-3)).all()
    assert np.abs(sn - d["sec_n"])[h].max() < 1e-4


@pytest.mark.parametrize("scene,W,H,spp", [("cornell_box", 40, 40, 48), ("generated_scene", 64, 36, 48)])
def test_render_matches_oracle_path_for_path(scene, W, H, spp):
    """same Philox stream -> the emulated kernel and the oracle trace the same paths; only ulp-level arithmetic differs"""
    objs, tex, sky_idx, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{scene}.json", W, H)
    earth, sky = imgio.read_png(pt.ASSETS + "/earth.png"), imgio.read_hdr(pt.ASSETS + "/skybox.hdr")
    O, E = orc.Oracle(objs), Emu(objs)
    assert O.add_texture(earth) == 1 and E.add_texture(earth) == 1
    if scene == "generated_scene":
        O.set_skybox(O.add_texture(sky))
        E.set_skybox(E.add_texture(sky))
    a, ra = O.render(cam, W, H, spp)
    b, rb = E.render(cam, W, H, spp)
    assert abs(int(ra) - int(rb)) <= 1e-4 * ra
    rel = np.abs(a - b)[..., :3] / (np.abs(a[..., :3]) + 1e-3 * spp)
    assert (rel.max(-1) > 1e-3).mean() < 0.01  # a few paths diverge at silhouettes; everything else agrees
    assert abs(a[..., :3].mean() / b[..., :3].mean() - 1) < 1e-3
    # accumulation semantics (ignoreHistory = false adds onto the buffer) and the sample partition
    c, _ = E.render(cam, W, H, spp // 2)
    c, _ = E.render(cam, W, H, spp - spp // 2, sample_offset=spp // 2, accum=c)
    assert np.allclose(b, c, rtol=1e-5, atol=1e-5)
    parts = [E.render(cam, W, H, (spp - r + 2) // 3, sample_offset=r, sample_stride=3)[0] for r in range(3)]
    assert np.allclose(sum(p[..., :3] for p in parts), b[..., :3], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("scene,W,H", [("cornell_box", 48, 48), ("generated_scene", 160, 90)])
def test_pixel_beam_lists_do_not_change_the_image(scene, W, H):
    """camera rays that take their leaves from the per-pixel beam list (beamLeaves, trace_device.cuh) find the same closest
    hits as the walk from the root: the image is bit-identical, the ray count equal, the lists short"""
    objs, tex, sky_idx, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{scene}.json", W, H)
    E = Emu(objs)
    E.add_texture(imgio.read_png(pt.ASSETS + "/earth.png"))
    if scene == "generated_scene":
        E.set_skybox(E.add_texture(imgio.read_hdr(pt.ASSETS + "/skybox.hdr")))
    a, ra = E.render(cam, W, H, 24)
    b, rb = E.render(cam, W, H, 24, beam=True)
    pixels, entries, overflows, longest = E.beam_stats
    assert ra == rb and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert pixels == W * H and overflows == 0 and longest <= 16 and entries > 0
    if scene == "generated_scene":
        # pixels so large that their beams reach more than 16 leaves: those fall back to the walk from the root
        a, ra = E.render(cam, 16, 9, 24)
        b, rb = E.render(cam, 16, 9, 24, beam=True)
        assert E.beam_stats[2] > 0 and ra == rb and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_quirks_match_oracle():
    """cube entered from inside reports t = tMin (Q2); sphere accepts its far root beyond tMax (Hittable.inl:152-157)"""
    objs = [pt.make_object("CUBE"), pt.make_object("SPHERE", position=(5, 0, 0)), pt.make_object("QUAD", position=(5, 0.5, 0), rotation_deg=(0, 0, 90))]
    O, E = orc.Oracle(objs), Emu(objs)
    o = np.array([[0.2, 0.1, 0.0], [5.0, 0.2, 0.1], [5.2, 0.0, 0.0]], np.float32)
    d = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]], np.float32)
    io, to, no = O.trace_rays(o, d)
    ie, te, ne = E.trace_rays(o, d)
    assert io[0] == 0 and to[0] == np.float32(0.001)
    assert np.array_equal(io, ie) and np.allclose(to, te, rtol=1e-6) and np.allclose(no, ne, atol=1e-6)


def test_compact_atan2_acos_accuracy():
    """the equirectangular lookups use a compact atan2 / acos (trace_device.cuh fastAtan2): far below the 1/256-texel
    resolution of the texture unit the reference samples with (2e-7 rad is 1e-4 texel of a 4096-wide map)"""
    E = Emu([pt.make_object("SPHERE")], 4)
    rng = np.random.default_rng(7)
    v = rng.normal(size=(200000, 2)).astype(np.float32)
    v = np.concatenate([v, [[0, 1], [0, -1], [1, 0], [-1, 0], [1e-20, 1], [1, 1e-20], [-1e-20, -1], [3, -3]]]).astype(np.float32)
    a, _ = E.fast_angles(v[:, 0], v[:, 1])
    ref = np.arctan2(v[:, 0].astype(np.float64), v[:, 1].astype(np.float64))
    err = np.abs(a - ref)
    err = np.minimum(err, 2 * np.pi - err)      # +pi and -pi are the same direction
    assert err.max() < 4e-7, err.max()
    c = np.clip(rng.uniform(-1, 1, 200000), -1, 1).astype(np.float32)
    c = np.concatenate([c, [-1, 1, 0, 0.9999999, -0.9999999]]).astype(np.float32)
    _, ac = E.fast_angles(np.zeros_like(c), c)
    # near |c| = 1 acos is ill-conditioned in fp32 (the library routine has the same sqrt(1 - c^2) sensitivity)
    assert np.abs(ac - np.arccos(c.astype(np.float64))).max() < 4e-4
    mid = np.abs(c) < 0.99
    assert np.abs(ac - np.arccos(c.astype(np.float64)))[mid].max() < 1e-6


def test_first_hit_path_of_the_render_kernel_at_full_size():
    """the render kernel's first-hit path for pixel-centre rays (option jitter = 0: pixel beam, closestHitWW, hot-path
    arithmetic) at 1920x1080, on the CPU: with and without beam lists it finds the same closest hits - also on a 10 k-object
    scene seen from afar, where the float32 quadratic is so noisy that a ray passing outside an object can 'hit' it and the
    beam path has to check the leaf's box like the walk does (BeamBox) - and on the bundled scene it agrees with the oracle
    (= the reference's host arithmetic) up to the pixels FMA contraction decides"""
    W, H = 1920, 1080
    for name in ("generated_scene", "synthetic_10000"):
        if name.startswith("synthetic_"):
            objs, cam = scenegen.synthetic_scene(int(name.split("_")[1]), W, H)
        else:
            objs, _, _, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{name}.json", W, H)
        E = Emu(objs)
        ib, tb = E.first_hit(cam, W, H, beam=True)
        iw, tw = E.first_hit(cam, W, H, beam=False)
        assert np.array_equal(ib, iw) and np.array_equal(tb.view(np.uint32), tw.view(np.uint32)), name
        if name == "generated_scene":
            io, to, _ = orc.Oracle(objs).primary_pass(cam, W, H)
            assert (ib != io).sum() <= 64 and (ib >= 0).mean() > 0.5
            same = (ib == io) & (io >= 0)
            assert np.quantile(np.abs(tb - to)[same] / to[same], 0.9999) < 1e-5
