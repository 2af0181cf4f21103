#!/usr/bin/env python3
"""Deterministic stand-ins for the two image assets the reference's scenes name but its checkout lacks
(/root/reference/.MISSING_LARGE_BLOBS: PathtracerCUDA/earth.png, PathtracerCUDA/skybox.hdr), plus compact copies
of the two bundled scene descriptions (data fixtures named by BASELINE.json's configs).

  assets/earth.png    1024x512 RGB8   procedural land/sea/ice albedo (value-noise continents)
  assets/skybox.hdr   1024x512 RGBE   procedural sky gradient + soft sun + ground, flat (non-RLE) Radiance file
  assets/scenes/*.json                json-minified scene files (float literals preserved)

Both the reference binaries (oracle/_ref) and this repo's renderer read the SAME files, so parity is unaffected by
their content.  Fixed seed; re-running reproduces the files byte for byte (numpy + PIL versions permitting).
"""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def value_noise(h, w, cells, rng):
    """periodic-in-x bilinear value noise"""
    g = rng.random((cells // 2 + 1, cells)).astype(np.float32)
    ys = np.linspace(0, cells // 2, h, endpoint=False)
    xs = np.linspace(0, cells, w, endpoint=False)
    y0 = np.floor(ys).astype(int); x0 = np.floor(xs).astype(int)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
    y1 = np.minimum(y0 + 1, cells // 2); x1 = (x0 + 1) % cells
    a = g[y0][:, x0]; b = g[y0][:, x1]; c = g[y1][:, x0]; d = g[y1][:, x1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_earth(path, w=1024, h=512):
    rng = np.random.default_rng(1984)
    n = sum(value_noise(h, w, c, rng) * a for c, a in ((4, 1.0), (8, 0.5), (16, 0.25), (32, 0.125), (64, 0.0625))) / 1.9375
    lat = np.abs(np.linspace(-1, 1, h))[:, None]
    land = n > 0.52
    sea = np.stack([0.02 + 0.05 * n, 0.10 + 0.25 * n, 0.35 + 0.4 * n], -1)
    veg = np.stack([0.15 + 0.5 * (n - 0.5) * 2 + 0.25 * lat, 0.45 - 0.2 * lat + 0.2 * (n - 0.5), 0.10 + 0.1 * lat + 0.0 * n], -1)
    img = np.where(land[..., None], veg, sea)
    ice = (lat + 0.15 * (n - 0.5)) > 0.86
    img = np.where(ice[..., None], np.array([0.92, 0.94, 0.97]), img)
    img8 = (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8)
    Image.fromarray(img8, "RGB").save(path, optimize=True)


def float_to_rgbe(rgb):
    m = rgb.max(-1)
    mant, exp = np.frexp(m)
    scale = np.where(m > 1e-32, mant * 256.0 / np.maximum(m, 1e-38), 0.0)
    out = np.zeros(rgb.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, exp + 128, 0).astype(np.uint8)
    return out


def make_sky(path, w=1024, h=512):
    v = (np.arange(h) + 0.5) / h            # 0 = +y (zenith), 1 = -y
    u = (np.arange(w) + 0.5) / w            # phi / 2pi
    theta = v[:, None] * np.pi
    phi = u[None, :] * 2 * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), np.sin(theta) * np.sin(phi)], -1)
    up = d[..., 1]
    horizon = np.array([0.85, 0.90, 1.00]); zenith = np.array([0.18, 0.38, 0.90]); ground = np.array([0.22, 0.20, 0.18])
    t = np.clip(up, 0, 1)[..., None] ** 0.45
    sky = horizon * (1 - t) + zenith * t
    img = np.where((up >= 0)[..., None], sky, ground * (1.0 + 0.5 * up[..., None]))
    sun_dir = np.array([0.45, 0.60, 0.66]); sun_dir /= np.linalg.norm(sun_dir)
    c = np.clip(d @ sun_dir, -1, 1)
    ang = np.arccos(c)
    sun = 24.0 * np.exp(-(ang / 0.07) ** 2) + 1.2 * np.exp(-(ang / 0.35) ** 2)
    img = img + sun[..., None] * np.array([1.0, 0.93, 0.80])
    rng = np.random.default_rng(7)
    clouds = value_noise(h, w, 16, rng) * 0.6 + value_noise(h, w, 32, rng) * 0.4
    cl = np.clip((clouds - 0.55) * 4, 0, 1) * np.clip(up * 3, 0, 1)
    img = img * (1 - 0.5 * cl[..., None]) + 0.9 * cl[..., None]
    rgbe = float_to_rgbe(img.astype(np.float32))
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# procedural stand-in (tools/make_assets.py)\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())


def minify_scene(src, dst):
    with open(src) as f:
        data = json.load(f)
    with open(dst, "w") as f:
        json.dump(data, f, separators=(",", ":"))
        f.write("\n")


def main():
    os.makedirs(os.path.join(ROOT, "assets", "scenes"), exist_ok=True)
    make_earth(os.path.join(ROOT, "assets", "earth.png"))
    make_sky(os.path.join(ROOT, "assets", "skybox.hdr"))
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/PathtracerCUDA"
    if os.path.isdir(ref):
        for name in ("cornell_box.json", "generated_scene.json"):
            minify_scene(os.path.join(ref, name), os.path.join(ROOT, "assets", "scenes", name))


if __name__ == "__main__":
    main()
