"""ctypes wrapper of the TEST-ONLY host emulation of the device functions (tests/emu/emu_kernels.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from pathtracercuda_b200.abi import object_array

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libemu_kernels.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


class Emu:
    def __init__(self, objects, max_leaf=4, builder=0):
        build()
        L = C.CDLL(SO)
        self.L = L
        L.emu_scene_create.restype = C.c_void_p
        L.emu_scene_create.argtypes = [C.c_size_t, C.c_void_p, C.c_uint32]
        L.emu_scene_create2.restype = C.c_void_p
        L.emu_scene_create2.argtypes = [C.c_size_t, C.c_void_p, C.c_uint32, C.c_int]
        L.emu_scene_destroy.argtypes = [C.c_void_p]
        L.emu_add_texture.restype = C.c_uint32
        L.emu_add_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.emu_set_skybox.argtypes = [C.c_void_p, C.c_uint32]
        L.emu_scene_info.argtypes = [C.c_void_p] * 4
        L.emu_scene_arrays.argtypes = [C.c_void_p] * 3
        L.emu_primary.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_trace_rays.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_render.restype = C.c_uint64
        L.emu_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p]
        self.n = len(objects)
        self._arr = object_array(objects)
        self.s = L.emu_scene_create2(self.n, C.byref(self._arr), max_leaf, builder)  # builder: scene_compile.h kBuilderSah / kBuilderLbvh
        assert self.s, "compileScene failed"
        self._keep = []

    def __del__(self):
        try:
            self.L.emu_scene_destroy(self.s)
        except Exception:
            pass

    def info(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.L.emu_scene_info(self.s, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def arrays(self):
        nodes, _, _ = self.info()
        nd = np.zeros((nodes, 16), np.uint32)
        pr = np.zeros((self.n, 16), np.uint32)
        self.L.emu_scene_arrays(self.s, _p(nd), _p(pr))
        return nd, pr

    def fast_angles(self, y, x):
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        a, c = np.zeros_like(y), np.zeros_like(y)
        self.L.emu_fast_angles.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.emu_fast_angles(y.size, _p(y), _p(x), _p(a), _p(c))
        return a, c

    def first_hit(self, cam, w, h, beam=True):
        """the production kernel's first-hit path (pixel beam + closestHitWW, hot-path arithmetic) for pixel-centre rays"""
        idx = np.zeros(w * h, np.int32)
        t = np.zeros(w * h, np.float32)
        self.L.emu_first_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        self.L.emu_first_hit(self.s, C.byref(cam), w, h, int(beam), _p(idx), _p(t))
        return idx, t

    def global_count(self):
        self.L.emu_global_count.restype = C.c_uint32
        self.L.emu_global_count.argtypes = [C.c_void_p]
        return int(self.L.emu_global_count(self.s))

    def add_texture(self, img):
        img = np.ascontiguousarray(img)
        is_hdr = img.dtype == np.float32
        return self.L.emu_add_texture(self.s, img.shape[1], img.shape[0], int(is_hdr), _p(img))

    def set_skybox(self, h):
        self.L.emu_set_skybox(self.s, h)

    def primary_pass(self, cam, w, h):
        idx = np.zeros(w * h, np.int32)
        t = np.zeros(w * h, np.float32)
        st = np.zeros(2, np.uint64)
        self.L.emu_primary(self.s, C.byref(cam), w, h, _p(idx), _p(t), _p(st))
        return idx, t, st

    def trace_rays(self, o, d, tmin=0.001):
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        n = o.shape[0]
        idx = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32)
        self.L.emu_trace_rays(self.s, n, _p(o), _p(d), tmin, _p(idx), _p(t), _p(nrm))
        return idx, t, nrm

    def render(self, cam, w, h, spp, seed=1984, sample_offset=0, sample_stride=1, accum=None, max_bounces=5, beam=False):
        """beam=True: camera rays take their leaves from the per-pixel beam lists (beamLeaves) instead of walking the tree;
        self.beam_stats = (pixels, list entries, overflows, longest list) afterwards"""
        add = accum is not None
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        self.L.emu_set_beam(int(beam))
        rays = self.L.emu_render(self.s, C.byref(cam), w, h, spp, seed, sample_offset, sample_stride, int(add), max_bounces, _p(accum))
        st = (C.c_uint64 * 4)()
        self.L.emu_beam_stats(st)
        self.beam_stats = tuple(int(v) for v in st)
        self.L.emu_set_beam(0)
        return accum, rays
