"""Multi-GPU rendering with one process per GPU: every rank renders its share into its own float4
accumulation buffer and the buffers are combined with ONE sum-reduce (torch.distributed: NCCL over NVLink on GPUs, gloo
in the CPU tests).  The Philox stream is keyed on (pixel, global sample index, bounce), so the work can be cut either way:

  * "pixels"  (default): rank r renders ALL samples of the pixels r, r + world, ... (row-major, so neighbouring pixels -
                which cost alike - go to different ranks) and leaves zeros elsewhere.  The reduced image is bit-identical to
                the single-GPU image, and a pixel's samples stay on one GPU, which the per-pixel machinery of long renders
                (beam walk, first-bounce stratification) needs to pay off.  This REPLACES the split SURVEY.md §8e / north_star
                propose; measured at 8 GPUs: 7.96x against 6.8x for the sample split.
  * "samples": rank r renders the global sample indices r, r + world, ... of every pixel (north_star's split).

(The same two partitions inside ONE process, with one host thread per GPU and ncclReduce, are in the C ABI:
pt_create_multi, include/pt_b200.h.)"""
import numpy as np


def partition_samples(spp, rank, world):
    """Interleaved partition of the global sample indices 0..spp-1: rank r takes r, r+world, r+2*world, ...
    Returns (sample_offset, sample_stride, count) - the values of the "sample_offset"/"sample_stride" options and the
    spp argument of pt_render for that rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    count = (spp - rank + world - 1) // world if spp > rank else 0
    return rank, world, count


def partition_pixels(rank, world):
    """Interleaved partition of the pixels (row-major index): rank r takes r, r+world, ...  Returns (pixel_offset,
    pixel_stride) - the values of the "pixel_offset"/"pixel_stride" options."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank, world


def configure_partition(tracer, spp, rank, world, partition="pixels"):
    """Set the tracer's partition options; returns the spp argument of pt_render for this rank.  Safe to call before
    every render: setting an unchanged "sample_offset" does not rewind a progressive render's sample cursor."""
    if partition == "pixels":
        off, stride = partition_pixels(rank, world)
        tracer.setOption("sample_stride", 1)
        tracer.setOption("sample_offset", 0)
        tracer.setOption("pixel_stride", stride)
        tracer.setOption("pixel_offset", off)
        tracer.setOption("alpha", 1.0)
        return spp
    if partition != "samples":
        raise ValueError("partition must be 'pixels' or 'samples'")
    off, stride, count = partition_samples(spp, rank, world)
    tracer.setOption("pixel_stride", 1)
    tracer.setOption("pixel_offset", 0)
    tracer.setOption("sample_stride", stride)
    tracer.setOption("sample_offset", off)
    tracer.setOption("alpha", 1.0 if rank == 0 else 0.0)  # every rank writes every pixel: only one of them the alpha of 1 (trace.cu:198)
    return count


def reduce_accumulation(accum_tensor, dst=0, group=None, all_ranks=False, out=None):
    """The single collective of the path: sum the (H, W, 4) float32 accumulation buffers over ranks.  With `out` the sum
    goes into that tensor and the rank's own buffer is left as it was - a progressive render can go on accumulating into it
    and be reduced again later (an in-place reduce would count the other ranks' earlier samples twice next time)."""
    import torch.distributed as dist
    t = accum_tensor
    if out is not None:
        out.copy_(accum_tensor)
        t = out
    if all_ranks:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return t


def attach_torch_accumulator(tracer, device):
    """Allocate the accumulation buffer as a torch tensor (so torch.distributed can reduce it in place) and hand its
    device pointer to the tracer (pt_set_accum_device_ptr)."""
    import torch
    t = torch.zeros((tracer.height, tracer.width, 4), dtype=torch.float32, device=device)
    tracer.setAccumDevicePtr(t.data_ptr())
    return t


class PartitionedRender:
    """One rank's side of a partitioned (optionally progressive) render.

        R = PartitionedRender(tracer, rank, world, "pixels", accum=attach_torch_accumulator(tracer, dev))
        R.render(cam, 1024, ignore_history=True); R.render(cam, 1024, ignore_history=False) ...
        image_sum = R.reduce(dst=0)            # separate tensor; the per-rank partial sums stay intact
        spp_total, frames = R.counts()         # what the sum has to be divided by (per-sample mean / reference Q1 normalisation)

    The partition is configured ONCE; a rank whose share of a call is empty (spp < world with the sample split) zeroes its
    buffer when the call restarts the accumulation instead of leaving stale samples in it."""

    def __init__(self, tracer, rank, world, partition="pixels", accum=None):
        self.tracer, self.rank, self.world, self.partition, self.accum = tracer, rank, world, partition, accum
        self.total_spp = 0     # samples per pixel of the WHOLE job so far (all ranks)
        self.frames = 0        # render calls so far (the reference's Q1 normalisation counts calls)
        self._reduced = None
        configure_partition(tracer, 0, rank, world, partition)

    def share(self, spp):
        return spp if self.partition == "pixels" else partition_samples(spp, self.rank, self.world)[2]

    def render(self, camera, spp, ignore_history=True):
        count = self.share(spp)
        if ignore_history:
            self.total_spp, self.frames = 0, 0
        if count == 0 and ignore_history and self.accum is not None:
            self.accum.zero_()  # pt_render(spp = 0) launches nothing: do not let an earlier render's sums into the reduce
        self.tracer.render(camera, count, ignore_history)
        self.total_spp += spp
        self.frames += 1
        return count

    def reduce(self, dst=0, group=None, all_ranks=False):
        import torch
        if self._reduced is None:
            self._reduced = torch.empty_like(self.accum)
        return reduce_accumulation(self.accum, dst=dst, group=group, all_ranks=all_ranks, out=self._reduced)

    def counts(self):
        return self.total_spp, self.frames


def render_partitioned(tracer, camera, spp, rank, world, ignore_history=True, partition="samples"):
    """One-shot form (kept for callers that render once): configure, render this rank's share.  Returns the spp rendered."""
    count = configure_partition(tracer, spp, rank, world, partition)
    tracer.render(camera, count, ignore_history)
    return count
