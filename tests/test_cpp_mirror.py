"""The C++ host-side mirror of the reference's module interfaces (include/coherence_b200.hpp): it builds with
g++ against the product library, reports a missing device as Failure (no CPU path), and on a GPU produces the
same numbers as the Python mirror for the same scene — both sit on the same C ABI."""
import os
import subprocess

import numpy as np
import pytest

from coherence_renderer_b200 import abi, scene as S

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "host_mirror")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cpp")])


def test_cpp_mirror_builds_and_reports_missing_device():
    _build()
    out = subprocess.run([EXE, "--no-gpu"], capture_output=True, text=True, timeout=120).stdout
    import torch

    if torch.cuda.is_available():
        assert "context created" in out
    else:
        assert out.startswith("Failure: coh_init") and "no CPU fallback" in out


@pytest.mark.gpu
def test_cpp_mirror_matches_python_path(ctx):
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300).stdout
    vals = {}
    for line in out.splitlines():
        toks = line.split()
        for k in range(len(toks) - 1):
            if toks[k + 1].isdigit():
                vals[toks[k]] = int(toks[k + 1])
    assert "negative box: Failure" in out and "empty group: Failure" in out and "unexpected" not in out
    W, H = 200, 160
    tri = [(20.3, 20.1), (120.7, 30.2), (60.2, 150.9)]
    edges = abi.host_edgelist_of_subpath(S.polygon_segments(tri))
    hs, hm = ctx.shapeminshape_of_edgelist(edges, 0)
    assert vals["shape_card"] == ctx.shape_card(hs) and vals["minshape_card"] == ctx.shape_card(hm)
    mx = ctx.shape_difference(hs, hm)
    assert vals["maxshape_card"] == ctx.shape_card(mx)
    assert vals["opacity_sum"] == int(ctx.polygon_opacity(edges, 0, mx).astype(np.int64).sum())
    b = S.SceneBuilder()
    b.polygon(tri, S.Fill.plain(S.dissolve(S.rgba8(200, 30, 30), 180)))
    b.filter("monochrome", [S.polygon_segments([(60.0, 40.0), (150.0, 45.0), (140.0, 120.0), (70.0, 110.0)])])
    b.group_begin(pretrans=int(0.55 * 255.0))
    b.polygon([(40.0, 40.0), (190.5, 60.5), (90.0, 150.0)], S.Fill.plain(S.rgba8(10, 200, 40)))
    b.cpg("subtraction", [S.polygon_segments([(10.0, 90.0), (180.0, 80.0), (170.0, 150.0), (30.0, 140.0)])],
          [S.polygon_segments([(60.0, 100.0), (120.0, 100.0), (120.0, 130.0), (60.0, 130.0)])],
          S.Fill.gradient((20.0, 20.0), (150.0, 120.0), True, False, S.rgba8(255, 0, 0), S.rgba8(0, 0, 255)))
    b.group_end()
    b.polygon([(100.0, 10.0), (190.0, 12.0), (185.0, 70.0), (105.0, 66.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 0, 0), 120)), convolve=("gaussian", 3))
    b.begin_background()
    b.rectangle(S.rgba8(211, 211, 211), 0.0, 0.0, float(W), float(H))
    objs, n, nbg, e, p = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    ctx.render_frame(sc, (0, 0, W, H), abi.COH_RENDER_RECORD_U)
    px = ctx.fb_read_rgba(0, 0, W, H).reshape(-1).astype(object)
    chk = sum(int(v) * (1 + i % 7) for i, v in enumerate(px)) % (1 << 64)
    assert vals["checksum"] == chk
    hu = ctx.render_uncovered()
    assert vals["uncovered_card"] == ctx.shape_card(hu)
    assert vals["sum"] == int(ctx.fb_read_rgb888(0, 0, W, H).astype(np.int64).sum())
    for h in (hs, hm, mx, hu):
        ctx.shape_free(h)
    ctx.scene_free(sc)
