"""Regenerates tests/golden/oracle_frames.json: digests of frames and covered-so-far sets rendered by the ORACLE for a
few fixed scenes (lion, layered random polygons with brush strokes, groups with PreTrans, gradients, CPG, filters).
They pin the oracle and the CUDA path against drifting together; they are not outputs of the OCaml reference
(parity unpinned, see oracle/).  usage: python tools/make_golden.py"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tests.golden_scenes import SCENES  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:24]


def main():
    from oracle import pyoracle as O

    out = {"_comment": "sha256 prefixes of oracle renders (RGBA8 words, row-major) and of the flat covered-so-far set; see tools/make_golden.py"}
    for name, (build, W, H, update) in SCENES.items():
        objs, n, nbg, e, p = build().arrays()
        img, u = O.render_frame(objs, n - nbg, nbg, e, p, update, want_u=True)
        out[name] = {"size": [W, H], "update": list(update), "frame": digest(img), "uncovered": digest(u.astype(np.int32)), "nonclear": int((img != 0).sum())}
        print(name, out[name])
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_frames.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
