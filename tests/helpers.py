"""shared helpers of the test-suite"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_objects(d):
    import ctypes as C
    from pathtracercuda_b200.abi import CameraDesc, ObjectDesc
    raw = d["objects"].tobytes()
    n = len(raw) // C.sizeof(ObjectDesc)
    arr = (ObjectDesc * n).from_buffer_copy(raw)
    objs = [arr[i] for i in range(n)]
    cam = CameraDesc.from_buffer_copy(d["camera"].tobytes()) if "camera" in d.files else None
    return objs, cam


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def tie_aware_index_check(idx_a, t_a, idx_b, t_b, hit_t_fn, rel=1e-5):
    """Indices must match exactly, except at genuine geometric ties: a mismatching pixel is accepted only if the other
    side's object, intersected on its own with the same ray, hits at the same t (within `rel`) - SURVEY.md Q7/Q8.
    hit_t_fn(pixel, object_index) -> t or None.  Returns (n_mismatch, n_ties)."""
    mism = np.nonzero(idx_a != idx_b)[0]
    ties = 0
    for p in mism:
        ok = False
        if idx_a[p] >= 0 and idx_b[p] >= 0:
            ta = hit_t_fn(int(p), int(idx_a[p]))
            tb = hit_t_fn(int(p), int(idx_b[p]))
            if ta is not None and tb is not None and abs(ta - tb) <= rel * max(abs(ta), abs(tb)):
                ok = True
        assert ok, f"pixel {p}: index {idx_a[p]} (t={t_a[p]}) vs {idx_b[p]} (t={t_b[p]}) is not a geometric tie"
        ties += 1
    return len(mism), ties
