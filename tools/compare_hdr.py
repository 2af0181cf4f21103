#!/usr/bin/env python3
"""usage: tools/compare_hdr.py a.hdr b.hdr - mean ratio and RMSE of two Radiance images"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import imgio
a = imgio.read_hdr(sys.argv[1])[..., :3]; b = imgio.read_hdr(sys.argv[2])[..., :3]
print("shape", a.shape, "mean a/b", float(a.mean() / b.mean()), "rmse", imgio.rmse(a, b))
