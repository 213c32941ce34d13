#!/usr/bin/env python
"""Scene DATA of a text page of the reference for the GPU box (where /root/reference does not exist): the painted paths
of /root/reference/mintext1.pdf as coherence_renderer_b200.pdf_import reads them (user-space coordinates as written in
the file, glyph outlines as lines and curves), written to coherence_renderer_b200/scenes/mintext1.json.
Like scenes/lion.json this is geometry, not reference source.  Run in the build container only.

    python tools/make_text_fixture.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coherence_renderer_b200 import pdf_import as P  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/mintext1.pdf"
pdf = P.PdfFile.open(src)
elts, box = P.graphic_of_page(pdf)
paths = []
for kind, fill, winding, subpaths, spec in P.scene_of_graphic(elts):
    assert kind == "fill"
    paths.append({"colour": int(fill.c0), "winding": winding,
                  "subpaths": [[[v for p in s[1:] for v in p] for s in sp] for sp in subpaths if sp]})
out = os.path.join(ROOT, "coherence_renderer_b200", "scenes", "mintext1.json")
json.dump({"source": "mintext1.pdf (johnwhitington/coherence-renderer), first page, painted paths in paint order", "mediabox": box, "paths": paths},
          open(out, "w"), separators=(",", ":"))
print(len(paths), "paths", sum(len(sp) for p in paths for sp in p["subpaths"]), "segments", os.path.getsize(out), "bytes")
