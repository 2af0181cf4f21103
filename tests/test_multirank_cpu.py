"""world_size-2 gloo test of the multi-GPU host logic: sample partition + the single sum-reduce of the accumulation
buffers (pathtracercuda_b200/distributed.py), with the CPU oracle standing in for the per-rank renderer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pathtracercuda_b200 as pt
from pathtracercuda_b200.distributed import PartitionedRender, configure_partition, partition_pixels, partition_samples, reduce_accumulation


def test_partition_covers_every_sample_once():
    for spp in (1, 2, 7, 8, 100, 4096):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                off, stride, count = partition_samples(spp, r, world)
                seen += [off + k * stride for k in range(count)]
            assert sorted(seen) == list(range(spp)), (spp, world)
    with pytest.raises(ValueError):
        partition_samples(8, 2, 2)


def test_pixel_partition_covers_every_pixel_once_and_sets_the_options():
    for world in (1, 2, 3, 8):
        owner = np.full(1000, -1)
        for r in range(world):
            off, stride = partition_pixels(r, world)
            assert (owner[off::stride] == -1).all()
            owner[off::stride] = r
        assert (owner >= 0).all() and np.bincount(owner).max() - np.bincount(owner).min() <= 1
    class Stub:
        def __init__(self): self.opts = {}
        def setOption(self, k, v): self.opts[k] = v
    t = Stub()
    assert configure_partition(t, 4096, 3, 8, "pixels") == 4096
    assert t.opts == {"sample_stride": 1, "sample_offset": 0, "pixel_stride": 8, "pixel_offset": 3, "alpha": 1.0}
    assert configure_partition(t, 4096, 3, 8, "samples") == 512
    assert t.opts == {"sample_stride": 8, "sample_offset": 3, "pixel_stride": 1, "pixel_offset": 0, "alpha": 0.0}
    with pytest.raises(ValueError):
        configure_partition(t, 8, 0, 2, "tiles")


class _OracleTracer:
    """stands in for pathtracercuda_b200.Pathtracer in the CPU tests: same option / render semantics (sample cursor that a
    repeated, unchanged "sample_offset" does not rewind; pt_render(spp = 0) launches nothing), the oracle does the rendering"""

    def __init__(self, O, cam, W, H):
        self.O, self.cam, self.W, self.H = O, cam, W, H
        self.opts = {"sample_offset": 0, "sample_stride": 1, "pixel_offset": 0, "pixel_stride": 1, "alpha": 1}
        self.cursor = 0
        self.accum = torch.zeros((H, W, 4), dtype=torch.float32)

    def setOption(self, k, v):
        if k == "sample_offset" and int(v) != self.opts[k]:
            self.cursor = int(v)
        self.opts[k] = v if k == "alpha" else int(v)

    def render(self, cam, spp, ignore_history):
        if ignore_history:
            self.cursor = self.opts["sample_offset"]
        if spp == 0:
            return
        a = self.accum.numpy()
        own = (np.arange(self.W * self.H) % self.opts["pixel_stride"] == self.opts["pixel_offset"]).reshape(self.H, self.W)
        img, _ = self.O.render(cam, self.W, self.H, spp, sample_offset=self.cursor, sample_stride=self.opts["sample_stride"], stratify=0)
        img = np.where(own[..., None], img, np.float32(0)).astype(np.float32)
        a[...] = img if ignore_history else a + img
        self.cursor += spp * self.opts["sample_stride"]


def _worker_progressive(rank, world, port, out_dir):
    """two-call progressive render in both partitions + a call in which one rank has nothing to do"""
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H = 24, 24
    objs, tex, sky, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/cornell_box.json", W, H)
    O = orc.Oracle(objs)
    res = {}
    for partition in ("pixels", "samples"):
        T = _OracleTracer(O, cam, W, H)
        T.accum += 123.0  # stale contents from "an earlier render"
        R = PartitionedRender(T, rank, world, partition, accum=T.accum)
        R.render(cam, 6, True)
        first = R.reduce(dst=0).clone()
        R.render(cam, 5, False)   # continues the global sample indices: 6 .. 10
        both = R.reduce(dst=0).clone()
        assert R.counts() == (11, 2)
        R.render(cam, 1, True)    # sample split: rank 1's share is empty - its buffer must go to zero, not keep 11 samples
        one = R.reduce(dst=0).clone()
        res[partition] = (first.numpy(), both.numpy(), one.numpy())
    if rank == 0:
        ref6, _ = O.render(cam, W, H, 6, stratify=0)
        ref11, _ = O.render(cam, W, H, 11, stratify=0)
        ref1, _ = O.render(cam, W, H, 1, stratify=0)
        np.savez(os.path.join(out_dir, "prog.npz"), ref6=ref6, ref11=ref11, ref1=ref1, **{f"{k}_{i}": v[i] for k, v in res.items() for i in range(3)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_progressive_render_and_empty_share(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_progressive, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    d = np.load(tmp_path / "prog.npz")
    for part in ("pixels", "samples"):
        for i, ref in enumerate(("ref6", "ref11", "ref1")):
            assert np.allclose(d[f"{part}_{i}"][..., :3], d[ref][..., :3], rtol=1e-5, atol=1e-5), (part, ref)


def _worker(rank, world, port, out_dir):
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, spp = 32, 32, 12
    objs, tex, sky, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/cornell_box.json", W, H)
    O = orc.Oracle(objs)
    off, stride, count = partition_samples(spp, rank, world)
    acc, rays = O.render(cam, W, H, count, sample_offset=off, sample_stride=stride)
    t = torch.from_numpy(acc)
    reduce_accumulation(t, dst=0)
    # pixel partition: all samples of this rank's pixels (the oracle renders the frame, the rank keeps what it owns - the
    # CUDA kernel only ever touches its own pixels), zeros elsewhere, the same reduce
    full, _ = O.render(cam, W, H, spp)
    poff, pstride = partition_pixels(rank, world)
    own = (np.arange(W * H) % pstride == poff).reshape(H, W)
    t2 = torch.from_numpy(np.where(own[..., None], full, np.float32(0)).astype(np.float32))
    reduce_accumulation(t2, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
        np.save(os.path.join(out_dir, "reduced_pixels.npy"), t2.numpy())
        np.save(os.path.join(out_dir, "full.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single_render(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    red, full = np.load(tmp_path / "reduced.npy"), np.load(tmp_path / "full.npy")
    # same sample set, different float summation order (per-rank partial sums, then the reduce)
    assert np.allclose(red[..., :3], full[..., :3], rtol=1e-5, atol=1e-5)
    assert np.all(red[..., 3] == 2.0)  # every rank wrote alpha = 1 (trace.cu:198); the reduce sums them
    # pixel partition: x + 0 = x, the reduced image is the single-process image bit for bit (alpha 1: one owner per pixel)
    assert np.array_equal(np.load(tmp_path / "reduced_pixels.npy").view(np.uint32), full.view(np.uint32))
