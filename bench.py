#!/usr/bin/env python3
"""bench.py — the headline measurement: path tracing throughput on generated_scene.json at 1920x1080 (BASELINE.json
config 3: 4096 spp) in Mrays/s, on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--spp S] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" = one complete render of the configuration: every pixel, `spp` samples, up to 5 path segments per sample
(ray = one closest-hit query, sample = one path).  With N GPUs the work is partitioned
across the ranks - by default the PIXELS (rank r renders pixels r, r+N, ... with all their samples: a pixel's samples stay
on one GPU, which the per-pixel beam walk and sample order of long renders need; the image is bit-identical to the
single-GPU one), with --partition samples the SAMPLES of every pixel (global Philox sample indices, so the sample set is
the same at any N) - and the float4 accumulation buffers are summed with ONE NCCL reduce inside the timed region: strong scaling.

  value   rays of the whole job / device time of the step (CUDA events around the trace kernel on its own stream +
          CUDA events around the reduce), max over ranks; scene and accumulation buffer resident in HBM
  e2e     same metric through the C ABI with HOST buffers, per step: pt_set_scene from the host object array (BVH build +
          upload), pt_render, pt_get_hdr (normalise + 33 MB device->host copy), wall clock
  --impl reference   the reference's OWN trace.cu rebuilt for sm_100 (oracle/_ref/ref_pt, unmodified program) on the same
          GPU, same scene / resolution / spp; time = its own "ms GPU time"; rays = its own count (oracle/_ref/ref_gpu).
          The reference has no CPU renderer; its __host__ __device__ math built with g++ (oracle/_ref/libref_host.so) is
          timed on the box's cores as `cpu_baseline` in both arms.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = "generated_scene"   # --scene: cornell_box | generated_scene | synthetic_<N> (BASELINE.json configs 2-4)
W, H = 1920, 1080
SCENE_DIR = None            # directory holding <SCENE>.json and the assets it names (set by resolve_scene)


def resolve_scene(args):
    """bundled scenes live under assets/scenes; synthetic_<N> is generated into a temporary directory (seed 1984)"""
    global SCENE, W, H, SCENE_DIR
    import pathtracercuda_b200 as pt
    SCENE, W, H = args.scene, args.width, args.height
    if SCENE.startswith("synthetic_"):
        import tempfile
        from pathtracercuda_b200 import scenegen
        SCENE_DIR = tempfile.mkdtemp(prefix="ptb_scene_")
        os.symlink(pt.ASSETS + "/skybox.hdr", SCENE_DIR + "/skybox.hdr")
        scenegen.write_synthetic_scene(f"{SCENE_DIR}/{SCENE}.json", int(SCENE.split("_")[1]))
        return f"{SCENE_DIR}/{SCENE}.json", SCENE_DIR
    SCENE_DIR = pt.ASSETS + "/scenes"
    return f"{SCENE_DIR}/{SCENE}.json", pt.ASSETS


def workload_label(spp):
    cfg = {"generated_scene": "BASELINE.json config 3", "cornell_box": "BASELINE.json config 2"}.get(SCENE, "BASELINE.json config 4")
    if SCENE == "generated_scene" and (W, H) == (3840, 2160):
        cfg = "BASELINE.json config 5"
    return f"{SCENE}.json {W}x{H} {spp} spp ({cfg}), stand-in earth.png/skybox.hdr (reference assets not in its checkout)"

# algorithmic FP32 operations per unit of work, counted from the reference source
OPS_NODE_BOX = 24.0          # one ray/box slab test (AABB.inl:22-44); a two-box node visit = 2 of these
OPS_PRIM_COMMON = 33.0       # world->local ray (Hittable.inl:92-98)
OPS_PRIM_SHAPE_AVG = 33.0    # mean of the per-shape parts over generated_scene's shape mix
OPS_HIT = 40.0               # accepted hit: normal -> world, normalize, point, face-forward
OPS_SHADE = 300.0            # GGX / LAMBERT_GGX sample (483 of 484 materials); Lambert is 140
OPS_MISS = 10.0
OPS_CAMERA = 25.0


def clock_sampler(stop, out, index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        out["sm_max_mhz"] = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        while not stop.is_set():
            out["samples"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, n in names.items():
                if r & bit:
                    out["reasons"].add(n)
            time.sleep(0.1)
    except Exception as e:  # noqa
        out["error"] = repr(e)


def cpu_baseline(seconds=12.0):
    """the reference's host math (libref_host.so: its own Hittable::hit / AABB::hit / Material::sample with XORWOW, full
    5-segment paths) on this box's cores; falls back to our oracle port when the reference build is absent"""
    import pathtracercuda_b200 as pt
    from oracle import imgio, orc
    cores = os.cpu_count() or 1
    w, h = 480, 270
    objs, tex, sky, cam = pt.parse_scene_py(f"{SCENE_DIR}/{SCENE}.json", w, h)
    skyimg = imgio.read_hdr(pt.ASSETS + "/skybox.hdr")
    has_sky = bool(sky) and tex[sky - 1] == "skybox.hdr"  # cornell_box names a file that does not exist: black environment
    if orc.have_ref_host():
        R = orc.RefHost()
        for o in objs:
            o.material.texture = 0
        R.set_scene(objs)
        R.set_camera(cam)
        if has_sky:
            R.set_sky(skyimg)
        run = lambda spp: R.render(w, h, spp)[1]
        kind = "reference"
    else:
        O = orc.Oracle(objs)
        O.add_texture(imgio.read_png(pt.ASSETS + "/earth.png"))
        if has_sky:
            O.set_skybox(O.add_texture(skyimg))
        run = lambda spp: O.render(cam, w, h, spp)[1]
        kind = "port"
    t0 = time.perf_counter()
    rays = run(2)
    rate = rays / (time.perf_counter() - t0)
    spp = int(max(2, min(4096, seconds * rate / (rays / 2))))
    t0 = time.perf_counter()
    rays = run(spp)
    dt = time.perf_counter() - t0
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"{SCENE} {w}x{h}, {spp} spp, full 5-segment paths ({rays} rays in {dt:.1f} s), OpenMP over scanlines"
                      + (", reference host build (base-colour texture taps compiled out by the reference itself)" if kind == "reference" else "")}


def run_reference(args):
    """--impl reference: the unmodified reference program on this box's GPU 0"""
    import pathtracercuda_b200 as pt
    from oracle import orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene_path, scene_cwd = resolve_scene(args)
    line = {"impl": "reference", "metric": "Mrays/s", "unit": "Mrays/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(args.spp), "spp": args.spp,
                       "l2": "not flushed: the reference re-launches its kernel every 8 spp, working set 56 KB scene + 133 MB state"}}
    cb = cpu_baseline()
    if not os.path.exists(orc.REF_PT):
        # no GPU build of the reference available: the CPU host build is the reference arm
        line.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb, "reference_device": f"host CPU x{cb['cores']}",
                     "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line))
        return 0
    objs, tex, sky, cam = pt.parse_scene_py(scene_path, W, H)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        cnt = orc.ref_gpu_count(objs, cam, W, H, 32, td)  # the reference's own rays per sample (its seeds, its slicing)
    rps = cnt["rays_per_sample"]
    times, walls = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        p = subprocess.run([orc.REF_PT, "-w", str(W), "-h", str(H), "-spp", str(args.spp), os.path.relpath(scene_path, scene_cwd)], cwd=scene_cwd, capture_output=True, text=True)
        wall = time.perf_counter() - t0
        ms = [float(l.split(" in ")[1].split(" ms")[0]) for l in p.stdout.splitlines() if l.startswith("Finished accumulating")]
        if p.returncode != 0 or not ms:
            print(json.dumps({"impl": "reference", "unavailable": "ref_pt failed: " + (p.stderr or p.stdout)[-200:].replace("\n", " ")}))
            return 0
        if i >= args.warmup:
            times.append(ms[0])
            walls.append(wall)
    ms = float(np.mean(times))
    rays = W * H * args.spp * rps
    line.update({"value": rays / ms / 1e3, "ms_per_step": ms, "samples_per_s": W * H * args.spp / ms * 1e3, "rays_per_sample": rps,
                 "reference_device": "B200: the reference's trace.cu / Pathtracer.cpp / main.cpp rebuilt unmodified for sm_100 (oracle/_ref/ref_pt); "
                                     "time = its own 'ms GPU time', rays from oracle/_ref/ref_gpu count",
                 "gpu_launches": (args.spp + 7) // 8 * args.steps, "cpu_baseline": cb,
                 "e2e": {"value": rays / (float(np.mean(walls)) * 1e3) / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                         "note": "whole process wall clock (CUDA init, scene load, 512 launches), no image written"}})
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--scene", default="generated_scene")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--partition", default="pixels", choices=["pixels", "samples"],
                    help="N > 1: what is split over the GPUs (pathtracercuda_b200/distributed.py); the single reduce is the same")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import pathtracercuda_b200 as pt
    from pathtracercuda_b200.distributed import attach_torch_accumulator, configure_partition, reduce_accumulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the trace path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE {world}"

    path, scene_cwd = resolve_scene(args)
    P = pt.Pathtracer(W, H, device=local)
    cam = P.loadSceneFile(path, cwd=scene_cwd)
    P.setOption("variant", args.variant)
    accum = attach_torch_accumulator(P, torch.device("cuda", local))
    count = configure_partition(P, args.spp, rank, world, args.partition)  # spp argument of pt_render on this rank
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """one render of the whole configuration: this rank's samples + the reduce.  returns (device ms, rays)"""
        flush.zero_()  # evict L2 between steps (outside the timed events)
        torch.cuda.synchronize()
        P.render(cam, count, True)  # blocking; timed by CUDA events on the library's own stream
        ms = P.getTiming()
        rays = P.stats().rays
        if dist:
            ev0.record()
            reduce_accumulation(accum, dst=0)
            ev1.record()
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
        return ms, rays

    for _ in range(args.warmup):
        step()
    clk = {"samples": [], "reasons": set()}
    stop = threading.Event()
    th = threading.Thread(target=clock_sampler, args=(stop, clk, local), daemon=True)
    th.start()
    barrier()
    t0 = time.perf_counter()
    ms_steps, rays_steps = [], []
    for _ in range(args.steps):
        ms, rays = step()
        ms_steps.append(ms)
        rays_steps.append(rays)
    barrier()
    wall = time.perf_counter() - t0
    stop.set()
    th.join(timeout=2)

    ms_t = torch.tensor(ms_steps, dtype=torch.float64, device=f"cuda:{local}")
    rays_t = torch.tensor(rays_steps, dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
    ms_step = float(ms_t.mean())
    rays_step = float(rays_t.mean())
    value = rays_step / ms_step / 1e3  # Mrays/s, whole job

    # ---- e2e: host buffers in, host image out, through the public API, wall clock ----
    objs, tex, sky, cam_py = pt.parse_scene_py(path, W, H)
    for o in objs:  # texture handles as loaded above: earth.png = 1 (skybox = 2)
        pass
    st0 = P.stats()
    h2d = int(st0.scene_bytes)
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for _ in range(args.steps):
        P.setScene(objs)              # host object array -> BVH build -> H2D of nodes/primitives/materials
        P.render(cam, count, True)
        e2e_rays += P.stats().rays
        if dist:
            reduce_accumulation(accum, dst=0)
            torch.cuda.synchronize()
        if rank == 0:
            img = P.getHDRImageData(copy=False)  # normalise on device + D2H of W*H*16 bytes into pinned host memory
            checksum = float(img[::97, ::89, :3].sum())
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_rays], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_t[0]) / e2e_wall / 1e6

    if rank == 0:
        # ---- roofline: FP32 issue.  Unit counts from one untimed counting launch. ----
        P.setOption("count_work", 1)
        P.render(cam, 8, True)
        s = P.stats()
        P.setOption("count_work", 0)
        per_ray = {"node_visits": s.node_visits / s.rays, "prim_tests": s.prim_tests / s.rays, "shades": s.shades / s.rays, "misses": s.misses / s.rays,
                   "camera_rays": s.samples / s.rays}
        ops_per_ray = (per_ray["node_visits"] * 2 * OPS_NODE_BOX + per_ray["prim_tests"] * (OPS_PRIM_COMMON + OPS_PRIM_SHAPE_AVG) + per_ray["shades"] * (OPS_HIT + OPS_SHADE)
                       + per_ray["misses"] * OPS_MISS + per_ray["camera_rays"] * OPS_CAMERA)
        sm_mhz = float(np.median(clk["samples"])) if clk["samples"] else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        f_mhz = sm_mhz or peaks.get("sm_max_mhz", 1965.0)
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        peak_tops = sms * 128 * f_mhz * 1e6 / 1e12          # FP32 lane-instruction slots per second (an FMA counts once)
        kernel_ms = float(np.mean(ms_steps))                # this rank's trace kernel (+ reduce) per launch
        achieved = rays_steps[0] * ops_per_ray / (kernel_ms * 1e-3) / 1e12
        prof = os.path.join(ROOT, "profiles", "r01b_trace_kernel_summary.json")
        traffic = None
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(args.spp),
                       "spp": args.spp, "spp_per_gpu": count,
                       "parallelism": ((f"pixel-partitioned x{world} (rank r: pixels r, r+{world}, ... with all {args.spp} samples; the other pixels of its buffer zero)" if args.partition == "pixels"
                                        else f"sample-partitioned x{world} (rank r: samples r, r+{world}, ... of every pixel)") + " + 1 NCCL sum-reduce of the float4 accumulation buffers") if world > 1 else "single GPU",
                       "l2": f"flushed between steps (256 MiB memset); scene ({st0.scene_bytes / 1e3:.0f} KB) " + ("is staged in shared memory" if st0.scene_in_smem else "is read through L1/L2") + f", output {W * H * 16 / 1e6:.0f} MB accumulation buffer",
                       "kernel_variant": args.variant},
            "samples_per_s": W * H * args.spp / ms_step * 1e3, "samples_per_s_per_gpu": W * H * args.spp / ms_step * 1e3 / world, "rays_per_sample": rays_step / (W * H * args.spp),
            "wall_ms_per_step": wall / args.steps * 1e3,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": W * H * 16,
                    "note": "per step: pt_set_scene (host BVH build + upload) + pt_render + pt_get_hdr (33 MB D2H to pinned memory), wall clock", "checksum": checksum},
            "gpu_launches": args.steps * world,
            "clocks": {"sm_mhz": sm_mhz, "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": sorted(clk["reasons"]), "n_samples": len(clk["samples"])},
            "roofline": {"bound": "fp32_issue", "achieved": achieved, "peak": peak_tops, "unit": "T FP32 lane-op/s", "frac": achieved / peak_tops, "traffic": traffic,
                         "note": "no dense contraction and a 64 KB on-chip scene: neither tensor nor HBM bound (DRAM traffic per launch in `traffic`, bytes). "
                                 "achieved = algorithmic FP32 ops (SURVEY.md §8d per-unit counts x unit counts of the plain walk from the root, from a counting launch with "
                                 "pixel beams off - the beams skip about half of those node visits, so this counts the algorithm's work, not executed instructions) / kernel time; "
                                 f"peak = {sms} SMs x 128 lanes x {f_mhz:.0f} MHz observed under load",
                         "ops_per_ray": ops_per_ray, "units_per_ray": per_ray},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    P.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
