#!/usr/bin/env python
"""Tokenise the uncompressed content stream of /root/reference/lion.pdf (obj 5: m/l/f/sc
only) into coherence_renderer_b200/scenes/lion.json.

Run in the build container only (the reference tree does not exist on the GPU box); the
JSON it writes is committed.  Geometry is scene DATA (PDF user-space coordinates and
DeviceRGB fill components as written in the file), not reference source code.
Paint order is kept (first path = painted first = back-most); the reference reverses it
when building the scene (examples.ml:174-180).
"""
import json
import os
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/lion.pdf"
data = open(src, "rb").read()
i = data.find(b"5 0 obj")
s = data.find(b"stream", i) + len(b"stream\n")
e = data.find(b"endstream", s)
toks = data[s:e].decode("latin1").split()

paths, stack, cur, rgb = [], [], None, [0.0, 0.0, 0.0]
num = re.compile(r"^-?(\d+\.?\d*|\.\d+)$")
for t in toks:
    if num.match(t):
        stack.append(t)
        continue
    if t == "sc":
        rgb = [float(x) for x in stack[-3:]]
    elif t == "m":
        cur = [[float(stack[-2]), float(stack[-1])]]
    elif t == "l":
        cur.append([float(stack[-2]), float(stack[-1])])
    elif t == "f":
        assert cur[0] == cur[-1], "subpath not explicitly closed"
        paths.append({"rgb": rgb, "winding": "nonzero", "subpaths": [cur]})
        cur = None
    stack = []
assert len(paths) == 132 and sum(len(p["subpaths"][0]) - 1 for p in paths) == 1930
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "coherence_renderer_b200", "scenes", "lion.json")
json.dump({"source": "lion.pdf obj 5 (johnwhitington/coherence-renderer)", "mediabox": [0, 0, 612, 792], "paths": paths}, open(out, "w"), separators=(",", ":"))
xs = [p[0] for q in paths for p in q["subpaths"][0]]
ys = [p[1] for q in paths for p in q["subpaths"][0]]
print("paths", len(paths), "bbox", min(xs), max(xs), min(ys), max(ys), "colours", len({tuple(p["rgb"]) for p in paths}))
