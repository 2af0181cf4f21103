#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: share of warp instructions, average active threads, stall samples.
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None
out = []
tot_inst = tot_thr = tot_samp = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] not in ("", "Line No") and r[2] == "-":
        try:
            inst, thr, samp = int(r[7]), int(r[8]), int(r[6])
        except ValueError:
            continue
        out.append((inst, thr, samp, cur, r[0], r[1].strip()[:100]))
        tot_inst += inst
        tot_thr += thr
        tot_samp += samp
print(f"total warp inst {tot_inst}  thread inst {tot_thr}  avg active threads {tot_thr / max(tot_inst, 1):.2f}  samples {tot_samp}")
out.sort(reverse=True)
for inst, thr, samp, f, ln, src in out[:top]:
    print(f"{100 * inst / tot_inst:5.2f}% inst  {100 * samp / max(tot_samp, 1):5.2f}% samp  thr {thr / max(inst, 1):5.1f}  {f}:{ln}  {src}")
