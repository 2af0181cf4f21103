"""`-m gpu`: the CUDA path, called through the C ABI (libpt_b200.so), against the oracle, the golden fixtures produced
by the reference's own GPU code, and size-independent properties at BASELINE.json's full sizes."""
import os
import subprocess

import numpy as np
import pytest

import pathtracercuda_b200 as pt
from pathtracercuda_b200 import scenegen
from oracle import imgio, orc
from tests.helpers import GOLDEN, ROOT, bits, golden_objects, load_golden

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "pathtracercuda_b200", "bin", "pathtracer_b200")
EARTH = None
SKY = None


def _assets():
    global EARTH, SKY
    if EARTH is None:
        EARTH, SKY = imgio.read_png(pt.ASSETS + "/earth.png"), imgio.read_hdr(pt.ASSETS + "/skybox.hdr")
    return EARTH, SKY


def _scene(name, W, H):
    if name.startswith("synthetic_"):
        objs, cam = scenegen.synthetic_scene(int(name.split("_")[1]), W, H)
        return objs, cam
    objs, tex, sky, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{name}.json", W, H)
    return objs, cam


# ---- gate 1: deterministic primary pass - index bit-exact, t within 1e-5 relative ---------------------------------------
@pytest.mark.parametrize("name", ["cornell_box", "generated_scene", "synthetic_1500"])
def test_primary_pass_vs_reference_gpu(name):
    """against the reference's own hitBVH run on a B200 (tests/golden/refgpu_primary_*.npz, tools/make_golden_gpu.py)"""
    g = load_golden(f"refgpu_primary_{name}")
    W, H = int(g["W"]), int(g["H"])
    objs, cam = _scene(name, W, H)
    with pt.Pathtracer(W, H) as P:
        P.setScene(objs)
        idx, t = P.primaryPass(cam)
    O = orc.Oracle(objs)
    mism = np.nonzero(idx != g["idx"])[0]
    for p in mism:  # a mismatch is only acceptable at an exact geometric tie (two objects at the same t, Q7/Q8)
        ray = O.camera_ray(cam, np.float32((p % W + 0.5) / W), np.float32((p // W + 0.5) / H))
        a, b = O.hit_object(int(idx[p]), ray[:3], ray[3:]), O.hit_object(int(g["idx"][p]), ray[:3], ray[3:])
        assert idx[p] >= 0 and g["idx"][p] >= 0 and a is not None and b is not None and abs(a[0] - b[0]) <= 1e-5 * a[0], p
    assert len(mism) <= 1e-5 * W * H + 4
    same = (idx == g["idx"]) & (idx >= 0)
    rel = np.abs(t - g["t"])[same] / g["t"][same]
    assert rel.max() <= 1e-5, rel.max()


@pytest.mark.parametrize("name", ["cornell_box", "generated_scene", "synthetic_1500"])
def test_primary_and_secondary_vs_oracle(name):
    """against the CPU oracle (= the reference's host math, pinned bit-exactly).  Host (no FMA) and device (FMA)
    arithmetic differ in the last bits (the reference GPU build differs from its own host build by up to 2.5e-5 in t),
    so silhouette pixels may flip: bounded, and everything else close."""
    d = load_golden(name)
    objs, cam = golden_objects(d)
    W, H = int(d["W"]), int(d["H"])
    with pt.Pathtracer(W, H) as P:
        P.setScene(objs)
        idx, t = P.primaryPass(cam)
        si, st, sn = P.traceRays(d["sec_o"], d["sec_d"])
    same = idx == d["primary_idx"]
    assert (~same).mean() < 2e-3
    h = same & (idx >= 0)
    rel = np.abs(t - d["primary_t"])[h] / d["primary_t"][h]
    assert np.quantile(rel, 0.99) < 1e-4 and rel.max() < 0.1  # FMA (device) vs no-FMA (host) rounding; the strict 1e-5 gate is vs the reference GPU above
    ssame = si == d["sec_idx"]
    assert ssame.mean() > 0.995
    h = ssame & (si >= 0)
    assert np.quantile(np.abs(st - d["sec_t"])[h] / np.maximum(d["sec_t"][h], 1.0), 0.99) < 1e-4  # tiny self-hit distances: absolute error
    assert np.quantile(np.abs(sn - d["sec_n"])[h].max(-1), 0.99) < 1e-3


def test_full_size_primary_properties():
    """1080p generated_scene: determinism, and agreement of the render path's first segment with the primary pass"""
    W, H = 1920, 1080
    objs, cam = _scene("generated_scene", W, H)
    with pt.Pathtracer(W, H) as P:
        P.setScene(objs)
        a = P.primaryPass(cam)
        b = P.primaryPass(cam)
        assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
        assert 0.55 < (a[0] >= 0).mean() < 0.60  # hit fraction 0.572
        # emissive-only check: paint every object with its own emission, 1 bounce, black sky -> image == f(index)
        for i, o in enumerate(objs):
            o.material.emissive[:] = [float(i + 1), 0.0, 0.0]
        P.setScene(objs)
        P.setOption("max_bounces", 1)
        P.render(cam, 4, True)
        img = P.getHDRMean()[..., 0].reshape(-1)
        # pixel-centre and jittered rays agree away from edges: compare on pixels whose 4 neighbours share the index
        idx2 = a[0].reshape(H, W)
        inner = np.ones((H, W), bool)
        for dy, dx in ((0, 1), (1, 0), (0, -1), (-1, 0), (1, 1), (-1, -1), (1, -1), (-1, 1)):
            inner &= np.roll(idx2, (dy, dx), (0, 1)) == idx2
        inner[0, :] = inner[-1, :] = inner[:, 0] = inner[:, -1] = False
        exp = (idx2 + 1).astype(np.float32)
        assert np.mean(np.abs(img.reshape(H, W)[inner] - exp[inner]) < 1e-3) > 0.999


# ---- gate 2: path tracing ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scene,W,H,spp", [("cornell_box", 64, 64, 64), ("generated_scene", 96, 54, 64), ("generated_scene", 80, 45, 320), ("cornell_box", 48, 48, 1100)])
def test_render_matches_oracle_path_for_path(scene, W, H, spp):
    """same Philox stream -> GPU and oracle trace the same paths; bit-level arithmetic differs, a few paths diverge.  From 128 spp
    both sides stratify the first scattering direction by the same integer rule (option "stratify", oracle/pt_oracle.c strataFor);
    1100 spp = 64 cells of 17 samples + 12 unstratified ones"""
    earth, sky = _assets()
    objs, cam = _scene(scene, W, H)
    O = orc.Oracle(objs)
    O.add_texture(earth)
    if scene == "generated_scene":
        O.set_skybox(O.add_texture(sky))
    a, ra = O.render(cam, W, H, spp)
    with pt.Pathtracer(W, H) as P:
        cam2 = P.loadSceneFile(f"{pt.ASSETS}/scenes/{scene}.json", cwd=pt.ASSETS)
        assert bytes(cam2) == bytes(cam)
        P.render(cam2, spp, True)
        b = P.getHDRMean() * spp
        st = P.stats()
    assert abs(int(st.rays) - int(ra)) <= 5e-4 * ra and st.samples == W * H * spp
    rel = np.abs(a - b)[..., :3] / (np.abs(a[..., :3]) + 1e-3 * spp)
    # a path that parts ways (a last-bit difference decides a hit or a lobe) moves its pixel's sum by ~1/spp of it: the more
    # samples a pixel has, the more of its paths do - so the per-pixel tolerance grows with spp, the share of pixels allowed
    # beyond it does not
    assert (rel.max(-1) > 1e-3 * max(1.0, spp / 64)).mean() < 0.04
    assert abs(a[..., :3].mean() / b[..., :3].mean() - 1) < 2e-3
    assert imgio.rmse(a / spp, b / spp)[0] < 0.02


@pytest.mark.parametrize("W,H,spp", [(96, 54, 32), (80, 45, 256)])
def test_env_importance_sampling_matches_oracle_path_for_path(W, H, spp):
    """option "env_is" (sky importance sampling + multiple importance sampling; not in the reference): the CUDA path against the
    oracle's restatement of the same estimator, same Philox counters -> same paths, same light samples, same shadow rays; both
    kernels that carry it (one pixel per lane below 64 spp, one pixel per warp above).  And against the plain render: same mean."""
    earth, sky = _assets()
    objs, cam = _scene("generated_scene", W, H)
    O = orc.Oracle(objs)
    O.add_texture(earth)
    O.set_skybox(O.add_texture(sky))
    plain, _ = O.render(cam, W, H, spp)
    O.set_env_is(1)
    a, ra = O.render(cam, W, H, spp)
    with pt.Pathtracer(W, H) as P:
        cam2 = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
        P.setOption("env_is", 1)
        P.render(cam2, spp, True)
        b = P.getHDRMean() * spp
        st = P.stats()
        P.render(cam2, spp, True)
        b2 = P.getHDRMean() * spp
        P.setOption("env_is", 0)
        P.render(cam2, spp, True)
        c = P.getHDRMean() * spp
        st0 = P.stats()
    assert np.array_equal(bits(b), bits(b2))                      # deterministic
    assert abs(int(st.rays) - int(ra)) <= 1e-3 * ra and st.rays > st0.rays and st.samples == W * H * spp
    rel = np.abs(a - b)[..., :3] / (np.abs(a[..., :3]) + 1e-3 * spp)
    assert (rel.max(-1) > 1e-3 * max(1.0, spp / 64)).mean() < 0.05
    assert abs(a[..., :3].mean() / b[..., :3].mean() - 1) < 2e-3
    assert imgio.rmse(a / spp, b / spp)[0] < 0.02
    # the same picture as without the option, less noise in it (against the oracle's plain render: an independent estimate)
    assert abs(b[..., :3].mean() / c[..., :3].mean() - 1) < 0.02
    assert not np.array_equal(bits(b), bits(c))


@pytest.mark.parametrize("scene", ["cornell_box", "generated_scene"])
def test_converged_image_within_reference_noise_floor(scene):
    """north_star gate: RMSE(ours, reference) <= 1.1 x RMSE(reference seed A, reference seed B) at 4096 spp on linear HDR,
    non-finite pixels masked and counted; the reference images come from the unmodified reference program on a B200."""
    g = load_golden(f"refpt_{scene}")
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    A, B = g["A"].astype(np.float32), g["B"].astype(np.float32)
    with pt.Pathtracer(W, H) as P:
        cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/{scene}.json", cwd=pt.ASSETS)
        P.render(cam, spp, True)
        img = P.getHDRMean()[..., :3]
        st = P.stats()
    floor, nfA = imgio.rmse(A, B)
    ours_a, nf1 = imgio.rmse(img, A)
    ours_b, nf2 = imgio.rmse(img, B)
    assert nf1 <= nfA + 8 and nf2 <= nfA + 8
    assert ours_a <= 1.1 * floor and ours_b <= 1.1 * floor, (ours_a, ours_b, floor)
    assert abs(st.rays / st.samples / float(g["rays_per_sample"]) - 1) < 5e-3
    fin = np.isfinite(A).all(-1)
    assert abs(img[fin].mean() / A[fin].mean() - 1) < 5e-3


def test_kernel_variants_bit_identical():
    """every scheduling variant of the trace kernel computes the same image: bit for bit among the one-pixel-per-lane
    kernels (counter-based RNG + per-pixel in-order accumulation make the result independent of which lane traces which
    path when) and among the one-pixel-per-warp kernels (same paths, summed as 32 per-lane partial sums + a butterfly, so
    only float reassociation separates the two families)"""
    def run(v, spp, **opts):
        with pt.Pathtracer(320, 180) as P:
            cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
            P.setOption("variant", v)
            for k, val in opts.items():
                P.setOption(k, val)
            P.render(cam, spp, True)
            P.render(cam, spp, False)
            return P.getHDRMean()
    imgs = {v: run(v, 16) for v in (1, 4, 5, 0)}
    for v in imgs:
        assert np.array_equal(bits(imgs[1]), bits(imgs[v])), v
    # one pixel per warp: 80 samples = two full rounds of 32 + a ragged one; deterministic run to run; the default for spp >= 128
    lane = run(4, 80)
    warp = [(v, run(v, 80)) for v in (8, 9, 10, 8)]
    for v, img in warp:
        assert np.array_equal(bits(warp[0][1]), bits(img)), v
    # camera passes / scattered passes hand the samples to the lanes in another order: its own sums, same from run to run
    split = run(12, 80)
    assert np.array_equal(bits(split), bits(run(12, 80))) and np.allclose(split, lane, rtol=2e-5, atol=1e-7)
    assert np.allclose(warp[0][1], lane, rtol=2e-5, atol=1e-7)
    assert np.allclose(run(8, 80, regen_low=1), lane, rtol=2e-5, atol=1e-7)
    assert np.allclose(run(8, 80, regen_low=32), lane, rtol=2e-5, atol=1e-7)
    assert np.array_equal(bits(run(0, 128)), bits(run(12, 128)))
    # pixel beams (camera rays take their leaves from the pixel's list instead of walking the tree): same closest hits
    assert np.array_equal(bits(run(8, 80, beam=1)), bits(run(8, 80, beam=0)))
    assert np.array_equal(bits(run(12, 80, beam=1)), bits(run(12, 80, beam=0)))
    # pixels so large that their beams reach more than 16 leaves (those fall back to the walk from the root)
    def coarse(beam):
        with pt.Pathtracer(16, 9) as P:
            cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
            P.setOption("variant", 12)
            P.setOption("beam", beam)
            P.render(cam, 96, True)
            return P.getHDRMean()
    assert np.array_equal(bits(coarse(1)), bits(coarse(0)))
    assert np.array_equal(bits(run(0, 256)), bits(run(12, 256, beam=0)))
    assert np.allclose(run(8, 5), run(4, 5), rtol=2e-5, atol=1e-7)  # fewer samples than lanes
    with pt.Pathtracer(320, 180) as P:  # shared-memory scene vs global-memory scene
        cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
        P.setOption("smem_scene", 0)
        P.render(cam, 32, True)
        assert P.stats().scene_in_smem == 0
        # a different template instantiation: nvcc may contract multiply-adds differently, so last-bit differences
        assert np.allclose(P.getHDRMean(), imgs[1], rtol=1e-4, atol=1e-5)


def test_accumulation_semantics_and_q1_normalisation():
    W, H = 128, 72
    with pt.Pathtracer(W, H) as P:
        cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/cornell_box.json", cwd=pt.ASSETS)
        P.render(cam, 24, True)
        one = P.getHDRMean()
        ref_norm = P.getHDRImageData()
        assert np.allclose(ref_norm, one * 24, rtol=1e-6)  # reference: sum / number of render() calls (Q1) = sum / 1
        assert np.all(ref_norm[..., 3] == 1.0)
        P.render(cam, 8, True)
        for _ in range(2):
            P.render(cam, 8, False)
        three = P.getHDRMean()
        assert np.allclose(three, one, rtol=2e-5, atol=1e-6)  # 3 x 8 samples continue the same global sample indices
        assert np.allclose(P.getHDRImageData()[..., :3], three[..., :3] * 8, rtol=1e-6)  # sum / 3 calls
        assert np.allclose(P.getHDRImageData()[..., 3], 1.0 / 3.0)  # alpha: kernel writes 1, host scales all four channels
        # the CLI's compat switch: one launch of 24 samples counted as ceil(24/8) = 3 reference calls
        P.setOption("frames_per_spp", 8)
        P.render(cam, 24, True)
        assert np.allclose(P.getHDRImageData()[..., :3], one[..., :3] * 8, rtol=1e-6)
        # tonemap (kernels/tonemap.cu) against the oracle's restatement on the same accumulation
        ldr = P.getImageData()
        O = orc.Oracle([pt.make_object("SPHERE")])
        exp = O.tonemap(P.getHDRMean() * 24, 3)
        assert (np.abs(ldr.astype(int) - exp.astype(int)) <= 1).all() and (ldr != exp).mean() < 0.01
        assert np.all(ldr[..., 3] == 255)
        # spp == 0 and the sample partition options
        P.setOption("frames_per_spp", 0)
        P.render(cam, 0, False)
        P.setOption("sample_stride", 2)
        P.setOption("sample_offset", 0)
        P.render(cam, 12, True)
        even = P.getHDRMean() * 12
        P.setOption("sample_offset", 1)
        P.render(cam, 12, True)
        odd = P.getHDRMean() * 12
        assert np.allclose((even + odd)[..., :3], one[..., :3] * 24, rtol=2e-5, atol=1e-5)


def test_edge_cases():
    with pt.Pathtracer(33, 17) as P:  # odd size, no scene: render is a no-op on a zeroed buffer
        cam = pt.make_camera((0, 0, 3), (0, 0, 0), 60, 33 / 17)
        P.setScene([])
        P.render(cam, 4, True)
        assert not P.getHDRImageData().any()
        idx, t = P.primaryPass(cam)
        assert (idx == -1).all()
        # a single object (root with an empty child), every shape, camera inside a cube (Q2)
        for shape in pt.SHAPES:
            P.setScene([pt.make_object(shape, rotation_deg=(30, 20, 10), emissive=(1, 1, 1))])
            P.setOption("max_bounces", 1)
            P.render(cam, 8, True)
            img = P.getHDRMean()
            assert np.isfinite(img).all() and img[..., :3].max() == 1.0
        P.setScene([pt.make_object("CUBE", scale=(10, 10, 10), emissive=(2, 2, 2))])
        idx, t = P.primaryPass(cam)
        assert (idx == 0).all() and (t == np.float32(0.001)).all()
        assert P.loadTexture("/nonexistent.png") == 0
        hs = [P.loadTextureMem(np.full((2, 2, 4), 128, np.uint8)) for _ in range(66)]
        assert hs[:64] == list(range(1, 65)) and hs[64:] == [0, 0]  # MAX_TEXTURE_COUNT 64
    with pytest.raises(pt.PtError):
        pt.Pathtracer(0, 10)


def test_textures_and_skybox():
    earth, sky = _assets()
    W, H = 96, 96
    objs = [pt.make_object("SPHERE", material="LAMBERT", texture=1), pt.make_object("QUAD", position=(0, -1, 0), scale=(3, 1, 3), texture=1),
            pt.make_object("CYLINDER", position=(2, 0, 0), scale=(0.5, 1, 0.5), texture=1), pt.make_object("DISK", position=(-2, 0, 0), rotation_deg=(90, 0, 0), texture=1)]
    cam = pt.make_camera((0, 1.5, 5), (0, 0, 0), 55, 1.0)
    O = orc.Oracle(objs)
    assert O.add_texture(earth) == 1
    O.set_skybox(O.add_texture(sky))
    a, _ = O.render(cam, W, H, 32)
    with pt.Pathtracer(W, H) as P:
        assert P.loadTexture(pt.ASSETS + "/earth.png") == 1 and P.loadTexture(pt.ASSETS + "/skybox.hdr") == 2
        P.setSkyboxTextureHandle(2)
        P.setScene(objs)
        P.render(cam, 32, True)
        b = P.getHDRMean() * 32
    rel = np.abs(a - b)[..., :3] / (np.abs(a[..., :3]) + 0.05)
    assert (rel.max(-1) > 2e-3).mean() < 0.03 and abs(a[..., :3].mean() / b[..., :3].mean() - 1) < 2e-3


def test_ragged_long_renders_are_deterministic():
    """the one-pixel-per-warp kernel with plain Philox draws (stratification off): sample counts that are not a multiple of
    the warp size, a second call accumulating onto the first, a sample partition (offset / stride) - the same bits from run
    to run, the same ray count however the lanes were filled"""
    def run(**opts):
        with pt.Pathtracer(160, 90) as P:
            cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
            P.setOption("stratify", 0)
            for k, v in opts.items():
                P.setOption(k, v)
            P.render(cam, 1056, True)
            P.render(cam, 1031, False)  # ragged second call
            return P.getHDRMean(), P.stats().rays
    a, ra = run()
    b, rb = run()
    c, rc = run(regen_low=8)  # lanes refilled in smaller groups: other summation order, same paths
    assert ra == rb == rc and np.array_equal(bits(a), bits(b)) and np.allclose(a, c, rtol=3e-5, atol=1e-7)
    a2, r2 = run(sample_offset=3, sample_stride=4)
    b2, r3 = run(sample_offset=3, sample_stride=4)
    assert r2 == r3 and np.array_equal(bits(a2), bits(b2)) and not np.allclose(a, a2, rtol=1e-3)


def test_first_bounce_stratification_is_unbiased_and_deterministic():
    """option "stratify" (default from 128 spp): the cell of the first scattering direction comes from the sample's index, the
    rest from Philox.  Same expectation as the plain draws: the two images agree within Monte Carlo noise (measured against
    two plain renders with different seeds), the ray count within its own noise; bit-identical from run to run; and a render
    too short for the default (64 spp) is untouched."""
    def run(strat, seed=1984, spp=2048, W=160, H=90):
        with pt.Pathtracer(W, H) as P:
            cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
            P.setOption("stratify", strat)
            P.setOption("seed", seed)
            P.render(cam, spp, True)
            return P.getHDRMean()[..., :3], P.stats().rays
    plain_a, ra = run(0)
    plain_b, rb = run(0, seed=7)
    strat_a, rs = run(1)
    strat_b, _ = run(-1)  # the default
    strat_c, _ = run(1, seed=7)
    assert np.array_equal(bits(strat_a), bits(strat_b)) and not np.array_equal(bits(strat_a), bits(plain_a))
    floor = imgio.rmse(plain_a, plain_b)[0]
    assert imgio.rmse(strat_a, plain_a)[0] <= 1.05 * floor and imgio.rmse(strat_a, plain_b)[0] <= 1.05 * floor
    assert imgio.rmse(strat_a, strat_c)[0] <= 1.02 * floor  # never more variance than the plain draws
    assert abs(strat_a.mean() / plain_a.mean() - 1) < 2e-3 and abs(rs / ra - 1) < 1e-3 and abs(rb / ra - 1) < 1e-3
    assert np.array_equal(bits(run(-1, spp=64)[0]), bits(run(0, spp=64)[0]))


def test_lbvh_builder_finds_the_same_hits(tmp_path):
    """option "bvh_builder" = 1 (Morton order + Karras' radix tree instead of the binned SAH): another tree over the same primitives,
    so the closest hit of every ray is the same - the IEEE primary pass agrees bit for bit, the production kernel's first hits agree,
    a path-traced image agrees up to the paths whose hit is decided by float rounding in the box tests"""
    W, H = 480, 270
    os.symlink(pt.ASSETS + "/skybox.hdr", str(tmp_path / "skybox.hdr"))
    scenegen.write_synthetic_scene(str(tmp_path / "scene.json"), 30000)
    res = {}
    for builder in (0, 1):
        with pt.Pathtracer(W, H) as P:
            P.setOption("bvh_builder", builder)
            cam = P.loadSceneFile(str(tmp_path / "scene.json"), cwd=str(tmp_path))
            idx, t = P.primaryPass(cam)
            P.setOption("jitter", 0)
            P.setOption("first_hit", 1)
            P.render(cam, 128, True)
            fi, ft = P.firstHit()
            P.setOption("jitter", 1)
            P.setOption("first_hit", 0)
            P.render(cam, 128, True)
            res[builder] = (idx, t, fi, ft, P.getHDRMean(), P.stats())
    a, b = res[0], res[1]
    assert b[5].bvh_nodes >= a[5].bvh_nodes and b[5].bvh_depth < 46      # one primitive per leaf: at least as many nodes
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
    assert (a[2] != b[2]).mean() < 1e-5 and np.allclose(a[3], b[3], rtol=1e-5, atol=1e-6)
    assert abs(int(a[5].rays) - int(b[5].rays)) <= 1e-3 * a[5].rays
    rel = np.abs(a[4] - b[4])[..., :3] / (np.abs(a[4][..., :3]) + 1e-3)
    share, means = float((rel.max(-1) > 1e-3).mean()), float(a[4][..., :3].mean() / b[4][..., :3].mean())
    print(f"lbvh vs sah: nodes {b[5].bvh_nodes} / {a[5].bvh_nodes}, depth {b[5].bvh_depth} / {a[5].bvh_depth}, first hits differing {(a[2] != b[2]).sum()}, pixels beyond 1e-3: {share:.4f}, mean ratio {means:.6f}")
    assert share < 0.02 and abs(means - 1) < 1e-3, (share, means)


def test_scene_sizes_around_the_shared_memory_opt_in_window():
    """scenes of 150 ... 3200 objects: node + primitive + material records (176 B per object) of 26 ... 560 KB.  Between 28 and 48 KB
    the scene fits the default dynamic limit only without the kernels' static shared memory (the opt-in has to be requested
    whenever dynamic shared memory is used); around 130-220 KB the scene fits but scene + shared-memory stack may not; beyond
    ~220 KB it stays in global memory.  Every size must render, finite, with the same rays per sample either way."""
    W, H = 96, 54
    for n in (150, 300, 400, 700, 1100, 1500, 3200):
        objs, cam = scenegen.synthetic_scene(n, W, H)
        res = []
        for smem in (1, 0):
            with pt.Pathtracer(W, H) as P:
                P.setOption("smem_scene", smem)
                P.setScene(objs)
                P.render(cam, 8, True)     # one pixel per lane
                st8 = P.stats()
                P.render(cam, 160, True)   # one pixel per warp, beams, stratified
                res.append((P.getHDRMean(), P.stats(), st8))
        (a, sa, sa8), (b, sb, sb8) = res
        assert np.isfinite(a).all() and sa.rays == sb.rays and sa8.rays == sb8.rays, n
        assert sb.scene_in_smem == 0 and sa.scene_in_smem == (1 if n <= 1100 else 0), (n, sa.scene_bytes)
        assert np.allclose(a, b, rtol=1e-4, atol=1e-5), n


@pytest.mark.parametrize("spp", [24, 320])
def test_pixel_partition_sums_to_the_single_gpu_image(spp):
    """multi-GPU pixel partition (options pixel_offset / pixel_stride): rank r renders pixels r, r+N, ... with all their
    samples and leaves zeros elsewhere, so the sum of the ranks' buffers IS the single-GPU image, bit for bit - here the
    three 'ranks' are three contexts on one GPU; also the accumulating second call"""
    W, H, N = 150, 70, 3
    def run(offset, stride):
        with pt.Pathtracer(W, H) as P:
            cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
            P.setOption("pixel_stride", stride)
            P.setOption("pixel_offset", offset)
            P.render(cam, spp, True)
            P.render(cam, spp, False)
            return P.getHDRMean().copy(), P.stats()
    full, st = run(0, 1)
    parts = [run(r, N) for r in range(N)]
    total = sum(p[0][..., :3] for p in parts)
    assert np.array_equal(bits(total), bits(full[..., :3]))
    assert sum(p[1].samples for p in parts) == st.samples == W * H * spp and sum(p[1].rays for p in parts) == st.rays
    for r, (img, _) in enumerate(parts):
        own = (np.arange(W * H) % N == r).reshape(H, W)
        assert not img[..., :3][~own].any() and img[..., :3][own].any()


def test_texture_unit_matches_software_filter():
    """texture taps through the texture unit (one TEX instruction, 1.8 fixed-point filter weights - the reference's own
    path) against the fp32 software filter over the packed texels (option tex_unit=0): same image up to the weights'
    1/256-texel quantisation"""
    imgs = []
    for unit in (1, 0):
        with pt.Pathtracer(240, 135) as P:
            cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
            P.setOption("tex_unit", unit)
            P.render(cam, 64, True)
            imgs.append(P.getHDRMean()[..., :3])
    hw, sw = imgs
    assert not np.array_equal(hw, sw)  # the option does switch paths
    rel = np.abs(hw - sw) / (np.abs(sw) + 0.05)
    assert (rel.max(-1) > 5e-3).mean() < 0.02 and abs(hw.mean() / sw.mean() - 1) < 1e-3


def test_cli_end_to_end(tmp_path):
    """the drop-in surface: same flags, same stdout lines, PNG / HDR files that decode to the API's images"""
    W, H, spp = 160, 90, 24
    png, hdr = str(tmp_path / "o.png"), str(tmp_path / "o.hdr")
    r = subprocess.run([CLI, "-w", str(W), "-h", str(H), "-spp", str(spp), "-o", png, "scenes/generated_scene.json"], cwd=pt.ASSETS, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[:9] == ["Beginning rendering in configuration:", f"Width: {W}", f"Height: {H}", f"Samples per Pixel: {spp}", "Window: 0", "Controls: 0",
                         "Output HDR: 0", f"Output Filepath: {png}", "Input Filepath: scenes/generated_scene.json"]
    assert "Accumulated 0 samples" in lines and any(l.startswith(f"Finished accumulating {spp} samples in ") and l.endswith(" ms GPU time") for l in lines)
    assert lines[-1] == f"Writing result to {png}"
    r = subprocess.run([CLI, "-w", str(W), "-h", str(H), "-spp", str(spp), "-ohdr", "-o", hdr, "scenes/generated_scene.json"], cwd=pt.ASSETS, capture_output=True, text=True)
    assert r.returncode == 0 and "Output HDR: 1" in r.stdout
    with pt.Pathtracer(W, H) as P:
        cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
        P.setOption("frames_per_spp", 8)
        P.render(cam, spp, True)
        ldr, hd = P.getImageData(), P.getHDRImageData()
    assert np.array_equal(imgio.read_png(png)[::-1], ldr)
    back = imgio.read_hdr(hdr)[::-1, :, :3]
    assert (np.abs(back - hd[..., :3]) <= hd[..., :3].max(-1, keepdims=True) / 128 + 1e-6).all()  # RGBE: 8 bits, shared exponent of the max channel
    r = subprocess.run([CLI, "missing.json"], cwd=pt.ASSETS, capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to open input file: missing.json" in r.stdout


def test_reference_program_agrees(tmp_path):
    """the unmodified reference program (oracle/_ref/ref_pt) and ours on the same scene, same box: means within noise"""
    if not os.path.exists(orc.REF_PT):
        pytest.skip("oracle/_ref/ref_pt not built")
    W, H, spp = 192, 108, 512
    out = str(tmp_path / "ref.hdr")
    r = subprocess.run([orc.REF_PT, "-w", str(W), "-h", str(H), "-spp", str(spp), "-ohdr", "-o", out, "scenes/generated_scene.json"], cwd=pt.ASSETS, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = imgio.read_hdr(out)[::-1, :, :3]
    with pt.Pathtracer(W, H) as P:
        cam = P.loadSceneFile(f"{pt.ASSETS}/scenes/generated_scene.json", cwd=pt.ASSETS)
        P.setOption("frames_per_spp", 8)
        P.render(cam, spp, True)
        ours = P.getHDRImageData()[..., :3]
    fin = np.isfinite(ref).all(-1) & np.isfinite(ours).all(-1)
    assert abs(ours[fin].mean() / ref[fin].mean() - 1) < 0.01
    assert imgio.rmse(ours, ref)[0] < 0.5 * ref[fin].mean()


def test_large_synthetic_scene_global_memory_path():
    """10k objects: the scene no longer fits in shared memory; the kernel reads it through the read-only path"""
    W, H = 320, 180
    objs, cam = scenegen.synthetic_scene(10000, W, H)
    O = orc.Oracle(objs)
    io, to, _ = O.primary_pass(cam, W, H)
    with pt.Pathtracer(W, H) as P:
        P.setScene(objs)
        idx, t = P.primaryPass(cam)
        P.render(cam, 8, True)
        st = P.stats()
        img = P.getHDRMean()
    assert st.scene_in_smem == 0 and st.bvh_nodes >= 5000
    assert (idx != io).mean() < 2e-3
    a, ra = O.render(cam, W, H, 8)
    # the host oracle (no FMA) re-hits curved surfaces slightly more often than any GPU build does: on quadric-only
    # scenes the reference's own GPU code traces 0.1-0.7 % fewer rays than its host build (measured) and so do we
    assert -1e-2 * ra < int(st.rays) - int(ra) < 2e-3 * ra
    assert np.isfinite(img).all()


def test_multi_gpu_context_through_the_c_abi(tmp_path):
    """pt_create_multi (two GPUs of one box, one host thread each inside pt_render): the same entry points as a single-device
    context.  Pixel partition - by peer-to-peer stores into the root's buffer and by one ncclReduce - is bit-identical to one
    GPU; the sample partition (north_star's split, always ncclReduce) agrees up to summation order; progressive calls keep
    accumulating; the getters normalise by the whole job's counts; the CLI's --gpus renders the same PNG."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    W, H, spp = 200, 120, 96
    scene = f"{pt.ASSETS}/scenes/generated_scene.json"

    def run(mask, partition=0, exchange=-1, stratify=-1):
        with pt.Pathtracer(W, H, device_mask=mask) as P:
            cam = P.loadSceneFile(scene, cwd=pt.ASSETS)
            if mask:
                P.setOption("partition", partition)
                P.setOption("exchange", exchange)
            P.setOption("stratify", stratify)
            P.render(cam, spp, True)
            first, st1, info = P.getHDRMean().copy(), P.stats(), P.multiInfo()
            P.render(cam, 40, False)
            both, frames = P.getHDRMean().copy(), P.getHDRImageData().copy()
            return first, both, frames, st1, info, P.getImageData().copy()
    one = run(0)
    p2p = run(3, 0, -1)
    red = run(3, 0, 0)
    assert one[4][0] == 1 and p2p[4][0] == 2 and red[4][0] == 2
    assert red[4][1] is False and red[4][3] > 0.0  # ncclReduce: its device time is reported on its own
    if p2p[4][1]:
        assert p2p[4][3] == 0.0                    # peer-to-peer stores: there is no exchange step
    for m in (p2p, red):
        assert np.array_equal(bits(m[0]), bits(one[0])) and np.array_equal(bits(m[1]), bits(one[1])) and np.array_equal(bits(m[2]), bits(one[2]))
        assert np.array_equal(m[5], one[5]) and m[3].rays == one[3].rays and m[3].samples == one[3].samples == W * H * spp
    # sample partition: each device stratifies its own launch, so compare with plain Philox draws (same sample set at any N)
    one_plain, smp = run(0, stratify=0), run(3, 1, 0, stratify=0)
    assert smp[4][1] is False
    assert np.allclose(smp[0], one_plain[0], rtol=3e-5, atol=1e-6) and np.allclose(smp[1], one_plain[1], rtol=3e-5, atol=1e-6)
    assert np.allclose(smp[2][..., :3], one_plain[2][..., :3], rtol=3e-5, atol=1e-6) and smp[3].rays == one_plain[3].rays
    assert np.allclose(smp[1][..., 3], one_plain[1][..., 3]) and np.allclose(smp[2][..., 3], one_plain[2][..., 3])  # alpha as on one GPU (only the first device writes the 1 of trace.cu:198)
    # the command line
    a, b = str(tmp_path / "one.png"), str(tmp_path / "two.png")
    for out, extra in ((a, []), (b, ["--gpus", "2", "--stats"])):
        r = subprocess.run([CLI, "-w", str(W), "-h", str(H), "-spp", "64", "-o", out] + extra + ["scenes/generated_scene.json"], cwd=pt.ASSETS, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    assert '"devices": 2' in r.stdout
    assert np.array_equal(imgio.read_png(a), imgio.read_png(b))


def test_render_kernels_inverse_length_is_the_ieee_one():
    """the camera-ray direction of the render kernels is normalised with the fast paths of the IEEE square root and division
    (trace_device.cuh invSqrtExact, no range checks, no slow-path calls): bit for bit 1.0f / sqrtf(x) as the reference's
    normalize() computes it - on the range a camera can produce and far beyond"""
    import ctypes as C
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(1.0, 6.0, 3_000_000), np.exp2(rng.uniform(-60, 60, 3_000_000)),
                        np.nextafter(np.float32(1.0), np.float32(2.0)) * np.ones(1), [1.0, 2.0, 3.0, 4.0, 0.25, 1e-20, 1e20]]).astype(np.float32)
    fast, ieee = np.zeros_like(x), np.zeros_like(x)
    with pt.Pathtracer(8, 8) as P:
        P.L.pt_debug_inv_sqrt.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        assert P.L.pt_debug_inv_sqrt(P.h, x.size, x.ctypes.data_as(C.c_void_p), fast.ctypes.data_as(C.c_void_p), ieee.ctypes.data_as(C.c_void_p)) == 0
    host = (np.float32(1.0) / np.sqrt(x)).astype(np.float32)
    assert np.array_equal(bits(ieee), bits(host))
    assert np.array_equal(bits(fast), bits(ieee)), f"{(bits(fast) != bits(ieee)).sum()} of {x.size} differ"


@pytest.mark.parametrize("scene,fmt", [("generated_scene", "hdr"), ("cornell_box", "png")])
def test_reference_front_end_on_our_back_end(scene, fmt, tmp_path):
    """oracle/_ref/ref_main_b200 = the reference's UNMODIFIED main.cpp + SceneLoader.cpp + Hittable.cpp + image writers, linked
    with integration/Pathtracer_b200.cpp (class Pathtracer over the C ABI) and libpt_b200.so instead of Pathtracer.cpp / BVH.cpp /
    kernels/*.cu.  It renders the bundled scenes - its own argument parsing, JSON loader, texture requests, 8-spp render loop,
    normalisation and file writers - and the image is the one our own command line writes for the same 8-spp slicing, bit for
    bit; and within noise of the reference program proper."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_main_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_main_b200 not built (oracle/Makefile needs /root/reference)")
    W, H, spp = 256, 144, 64
    outs = {}
    for tag, cmd in (("shim", [exe]), ("cli", [CLI, "--slice", "8"]), ("ref", [orc.REF_PT])):
        if tag == "ref" and not os.path.exists(orc.REF_PT):
            continue
        out = str(tmp_path / f"{tag}.{fmt}")
        r = subprocess.run(cmd + ["-w", str(W), "-h", str(H), "-spp", str(spp)] + (["-ohdr"] if fmt == "hdr" else []) + ["-o", out, f"scenes/{scene}.json"],
                           cwd=pt.ASSETS, capture_output=True, text=True)
        assert r.returncode == 0 and os.path.exists(out), (tag, r.stdout[-300:], r.stderr[-300:])
        assert f"Finished accumulating {spp} samples in " in r.stdout
        outs[tag] = imgio.read_hdr(out) if fmt == "hdr" else imgio.read_png(out)
    assert np.array_equal(outs["shim"], outs["cli"])
    if "ref" in outs:
        a, b = outs["shim"][..., :3].astype(np.float64), outs["ref"][..., :3].astype(np.float64)
        assert abs(a.mean() / b.mean() - 1) < 0.02
