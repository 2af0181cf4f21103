"""Option "env_is" (north_star item 5): importance sampling of the sky with multiple importance sampling.
Not in the reference - what has to hold is that it integrates the SAME measure (the image converges to the reference
estimator's) with less variance.  CPU part: the distribution tables (product == oracle restatement == an independent numpy
computation), the sampler against its own pdf, and the estimator in the oracle (unbiased against the plain one, lower RMSE).
The CUDA path is checked against the oracle path for path in tests/test_gpu_parity.py and at 1080p in test_gpu_baseline_sizes.py."""
import numpy as np
import pytest

import pathtracercuda_b200 as pt
from oracle import imgio, orc


def _sky():
    return imgio.read_hdr(pt.ASSETS + "/skybox.hdr")


def _oracle(scene, W, H):
    objs, tex, sky, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{scene}.json", W, H)
    O = orc.Oracle(objs)
    import os
    handles = []
    for t in tex:  # (a file that does not exist gets handle 0, as in the loader: cornell_box names a sky that is not there)
        f = f"{pt.ASSETS}/{t}"
        handles.append(0 if not os.path.exists(f) else O.add_texture(imgio.read_hdr(f) if t.endswith(".hdr") else imgio.read_png(f)))
    O.set_skybox(handles[sky - 1] if sky else 0)
    return O, cam


def test_distribution_tables():
    sky = _sky()
    cols, rows, q, alias, dens = pt.env_distribution(sky)
    assert (cols, rows) == (512, 256)
    # the oracle's restatement gives the same tables bit for bit
    O = orc.Oracle([pt.make_object("SPHERE")])
    O.set_skybox(O.add_texture(sky))
    O.set_env_is(1)
    c2, r2, q2, a2, d2 = O.env_tables()
    assert (c2, r2) == (cols, rows) and np.array_equal(q.view(np.uint32), q2.view(np.uint32)) and np.array_equal(alias, a2) and np.array_equal(dens.view(np.uint32), d2.view(np.uint32))
    # independent statement of what the tables must encode: P(cell) ~ sum of luminance x sin(theta) over the cell's texels (+ floor)
    H, W = sky.shape[:2]
    lum = sky[..., :3].astype(np.float64) @ np.array([0.2126, 0.7152, 0.0722])
    w = lum * np.sin(np.pi * (np.arange(H) + 0.5) / H)[:, None]
    w = w.reshape(rows, H // rows, cols, W // cols).sum((1, 3)).ravel()
    w += 1e-4 * w.sum() / w.size
    P = w / w.sum()
    n = P.size
    assert np.allclose(dens.astype(np.float64) * 2 * np.pi ** 2 / n, P, rtol=1e-6)
    # the alias table realises exactly these probabilities: P_i = (q_i + sum over j with alias_j = i of (1 - q_j)) / n
    back = q.astype(np.float64).copy()
    np.add.at(back, alias, 1.0 - q.astype(np.float64))
    assert np.allclose(back / n, P, rtol=2e-5, atol=1e-12)
    assert (q >= 0).all() and (q <= 1).all() and (alias < n).all()
    # an LDR map and a size that is not a multiple of the grid
    rng = np.random.default_rng(3)
    ldr = rng.integers(0, 256, (300, 700, 4), dtype=np.uint8)
    c3, r3, q3, a3, d3 = pt.env_distribution(ldr)
    assert c3 == 350 and r3 == 150 and abs(d3.astype(np.float64).sum() * 2 * np.pi ** 2 / (c3 * r3) - 1) < 1e-5


def test_sampler_follows_its_pdf():
    O = orc.Oracle([pt.make_object("SPHERE")])
    sky = _sky()
    O.set_skybox(O.add_texture(sky))
    O.set_env_is(1)
    cols, rows, q, alias, dens = O.env_tables()
    rng = np.random.default_rng(11)
    n = 120000
    s = O.env_sample(rng.integers(0, 2 ** 32, (n, 3), dtype=np.uint64).astype(np.uint32))
    d, u, v, pdf = s[:, :3], s[:, 3], s[:, 4], s[:, 5]
    assert np.allclose(np.linalg.norm(d, axis=1), 1, atol=1e-5)
    # direction <-> lookup coordinates: the convention of the sky lookup (trace.cu:123-127)
    theta, phi = np.arccos(np.clip(d[:, 1].astype(np.float64), -1, 1)), np.arctan2(d[:, 2].astype(np.float64), d[:, 0].astype(np.float64))
    pole = np.sin(theta) < 0.05  # (acos of a float32 cosine is ill-conditioned there)
    assert np.abs(theta / np.pi - v)[~pole].max() < 2e-6 and np.abs(theta / np.pi - v).max() < 1e-3
    du = (phi / (2 * np.pi) - u + 0.5) % 1.0 - 0.5
    assert np.abs(du[~pole]).max() < 2e-6
    # the pdf it reports is density / sin(theta) of the cell it landed in
    cell = np.minimum((v * rows).astype(int), rows - 1) * cols + np.minimum((u * cols).astype(int), cols - 1)
    assert np.allclose(pdf, dens[cell] / np.maximum(np.sin(np.pi * v.astype(np.float64)), 1e-6), rtol=1e-4)
    # and the samples follow it: counts in coarse blocks of 32 x 32 cells against the blocks' probabilities
    P = dens.astype(np.float64) * 2 * np.pi ** 2 / (cols * rows)
    blocks = P.reshape(rows // 32, 32, cols // 32, 32).sum((1, 3)).ravel()
    got = np.bincount((cell // cols // 32) * (cols // 32) + (cell % cols) // 32, minlength=blocks.size) / n
    sigma = np.sqrt(blocks * (1 - blocks) / n)
    assert (np.abs(got - blocks) < 5 * sigma + 1e-5).all()
    # the integral of (anything / pdf) over the samples estimates the integral over the sphere: the solid angle itself
    assert abs(np.mean(1.0 / pdf.astype(np.float64)) / (4 * np.pi) - 1) < 0.02


def test_estimator_same_measure_less_variance():
    """the oracle's statement of the estimator on the bundled scene: same expectation as the plain estimator (means agree far
    inside the noise), markedly lower error against a converged plain render at equal sample count"""
    W, H, spp = 64, 36, 256
    O, cam = _oracle("generated_scene", W, H)
    plain, rp = O.render(cam, W, H, spp, stratify=0)
    ref, _ = O.render(cam, W, H, 4096, seed=4242, stratify=0)
    O.set_env_is(1)
    mis, rm = O.render(cam, W, H, spp, stratify=0)
    mis2, _ = O.render(cam, W, H, spp, seed=99, stratify=0)
    O.set_env_is(0)
    plain, mis, mis2, ref = plain[..., :3] / spp, mis[..., :3] / spp, mis2[..., :3] / spp, ref[..., :3] / 4096
    assert rm > rp  # one more ray per scattering vertex with something to gain
    e_plain, e_mis = np.sqrt(((plain - ref) ** 2).mean()), np.sqrt(((mis - ref) ** 2).mean())
    assert e_mis < 0.75 * e_plain, (e_mis, e_plain)
    # means: |difference| against the spread two independent env_is renders show
    spread = abs(mis.mean() - mis2.mean())
    assert abs(mis.mean() / ref.mean() - 1) < 5e-3 and abs(0.5 * (mis.mean() + mis2.mean()) - ref.mean()) < 4 * spread + 2e-3 * ref.mean()
    # a scene without a sky is untouched by the option
    O2, cam2 = _oracle("cornell_box", 32, 32)
    a, ra = O2.render(cam2, 32, 32, 16)
    O2.set_env_is(1)
    b, rb = O2.render(cam2, 32, 32, 16)
    assert ra == rb and np.array_equal(a, b)
