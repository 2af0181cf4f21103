#!/bin/bash
# round-2 GPU run 36: process start-up of the binary (create phase, whole-process wall clock) before / after cudaDeviceGetAttribute
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY' | tee gpurun_out/r2_run36.txt
import subprocess, time, re, os
root = os.getcwd()
for i in range(4):
    for b in ("build/exp/head/pathtracercuda_b200/bin/pathtracer_b200", "pathtracercuda_b200/bin/pathtracer_b200"):
        t0 = time.perf_counter()
        p = subprocess.run([os.path.join(root, b), "-w", "1920", "-h", "1080", "-spp", "64", "-ohdr", "-o", "/tmp/o.hdr", "--stats", "scenes/generated_scene.json"], cwd="assets", capture_output=True, text=True)
        wall = time.perf_counter() - t0
        m = re.search(r'"host_ms": \{[^}]*\}', p.stdout)
        print(f"{wall:.3f} s wall  {m.group(0) if m else p.stdout[-200:] + p.stderr[-200:]}  <- {b}", flush=True)
PY
