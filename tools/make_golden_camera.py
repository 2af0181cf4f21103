#!/usr/bin/env python3
"""tests/golden/camera_moves.npz: the reference's own Camera::rotate / Camera::translate (Camera.inl:30-52, through
oracle/_ref/libref_host.so) applied to the cameras of the bundled scenes - the 23 floats of the reference's Camera after
every move.  Run in the build container (needs /root/reference for the harness build); the fixture is committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracercuda_b200 as pt
from oracle import orc

R = orc.RefHost()
rng = np.random.default_rng(424242)
starts, moves, states = [], [], []
for scene, (W, H) in {"cornell_box": (256, 256), "generated_scene": (1920, 1080)}.items():
    objs, tex, sky, cam = pt.parse_scene_py(f"{pt.ASSETS}/scenes/{scene}.json", W, H)
    R.set_camera(cam)
    starts.append(np.frombuffer(bytes(cam), np.uint8).copy())
    mv, st = [], [R.get_camera()]
    for k in range(24):
        if k % 2 == 0:
            m = np.array([0, rng.uniform(-0.4, 0.4), rng.uniform(-0.8, 0.8), rng.uniform(-1, 1)], np.float32)  # pitch, yaw, roll (ignored)
            R.camera_rotate(float(m[1]), float(m[2]), float(m[3]))
        else:
            m = np.array([1, *rng.uniform(-2, 2, 3)], np.float32)
            R.camera_translate(float(m[1]), float(m[2]), float(m[3]))
        mv.append(m)
        st.append(R.get_camera())
    moves.append(np.array(mv))
    states.append(np.array(st))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "camera_moves.npz"), start=np.array(starts), moves=np.array(moves), states=np.array(states))
print("camera moves", np.array(states).shape)
