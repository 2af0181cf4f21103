"""Independent (numpy/PIL) image decoders used by the TESTS to cross-check the product's own C++ readers/writers
(pathtracercuda_b200/csrc/image_io.cpp).  Decode rules follow the reference's vendored stb_image 2.26
(reference: PathtracerCUDA/src/stb_image.h:7036-7062 — value = mantissa * 2^(e-136), no +0.5, alpha = 1)."""
import numpy as np


def read_hdr(path):
    """Radiance RGBE -> float32 RGBA (H, W, 4), top row first.  Handles flat and new-style RLE scanlines."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    lines = []
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if line == b"":
            break
        lines.append(line)
    assert lines[0] in (b"#?RADIANCE", b"#?RGBE"), lines[0]
    end = data.index(b"\n", pos)
    dims = data[pos:end].split()
    pos = end + 1
    assert dims[0] == b"-Y" and dims[2] == b"+X"
    h, w = int(dims[1]), int(dims[3])
    buf = np.frombuffer(data, np.uint8, offset=pos)
    rgbe = np.zeros((h, w, 4), np.uint8)
    if w < 8 or w >= 32768 or not (buf[0] == 2 and buf[1] == 2 and (buf[2] & 0x80) == 0):
        rgbe[:] = buf[: h * w * 4].reshape(h, w, 4)
    else:
        p = 0
        for y in range(h):
            assert buf[p] == 2 and buf[p + 1] == 2 and ((int(buf[p + 2]) << 8) | int(buf[p + 3])) == w
            p += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = int(buf[p]); p += 1
                    if n > 128:
                        n -= 128
                        rgbe[y, x:x + n, c] = buf[p]; p += 1
                    else:
                        rgbe[y, x:x + n, c] = buf[p:p + n]; p += n
                    x += n
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e != 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    out = np.ones((h, w, 4), np.float32)
    out[..., :3] = rgbe[..., :3].astype(np.float32) * scale[..., None]
    return out


def read_png(path):
    """PNG -> uint8 RGBA (H, W, 4), top row first (what stbi_load(..., 4) returns)."""
    from PIL import Image
    return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGBA"), np.uint8))


def rmse(a, b, mask_nonfinite=True):
    """RMSE over RGB of two (H, W, >=3) float images; non-finite pixels in either are masked and counted."""
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    ok = np.isfinite(a).all(-1) & np.isfinite(b).all(-1)
    d = (a - b)[ok]
    return float(np.sqrt(np.mean(d * d))), int((~ok).sum())
