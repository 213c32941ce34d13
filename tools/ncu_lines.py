#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line (instructions executed, stall samples).

  python tools/ncu_lines.py report.ncu-rep [top] [kernel-name-substring] [inst|samp]
"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = sys.argv[3] if len(sys.argv) > 3 else None
key = sys.argv[4] if len(sys.argv) > 4 else "samp"
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]
if kern: cmd += ["--kernel-name", kern]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg, cur = {}, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        try:
            k = (cur, int(r[0])); v = agg.get(k, (0, 0, 0, ""))
            agg[k] = (v[0] + int(r[7]), v[1] + int(r[6]), v[2] + int(r[8]), r[1].strip()[:100])
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot/1e6:.1f}M, stall samples {tots}")
ki = 0 if key == "inst" else 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][ki])[:top]:
    print(f"{k[0]}:{k[1]:4d} inst={v[0]/1e6:7.2f}M ({100*v[0]/tot:4.1f}%) samp={100*v[1]/max(tots,1):4.1f}% thr/inst={v[2]/max(v[0],1):4.1f} | {v[3]}")
