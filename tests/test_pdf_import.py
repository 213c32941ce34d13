"""N4 — PDF import (coherence_renderer_b200/pdf_import.py): Pdfgraphics.graphic_of_page's path side (pdfgraphics.ml:727-1245)
and Render.scene_of_graphic (render.ml:1476-1565) over a minimal PDF reader.

CPU: PDFs written by this file exercise every operator the importer knows and the places where the reference departs from the
PDF specification (kept on purpose: the scenes the raster path renders are the ones the reference's reader produces); the
host route of the stroker's edges against the oracle; and — only where /root/reference exists (this container, not the GPU
box) — the reference's own fixture PDFs: lion.pdf must give exactly the committed lion fixture, every file must import.
GPU: an imported page (fills, strokes, transparency, even-odd, a form XObject) rendered against the oracle.
"""
import glob
import os
import zlib

import numpy as np
import pytest

from coherence_renderer_b200 import abi, pdf_import as P, scene as S

REF = "/root/reference"


def make_pdf(content, extgstate="", xobjects="", colorspaces="", extra_objects=b"", compress=True, inherit=True):
    """A one-page PDF around a content stream (no cross-reference table: the reader finds objects by their headers)."""
    body = zlib.compress(content) if compress else content
    filt = b"/Filter /FlateDecode " if compress else b""
    res = ("<< /ExtGState << %s >> /XObject << %s >> /ColorSpace << %s >> >>" % (extgstate, xobjects, colorspaces)).encode()
    page_res = b"" if inherit else b"/Resources " + res
    tree_res = b"/Resources " + res if inherit else b""
    return (b"%PDF-1.4\n1 0 obj\n<< /Type /Catalog /Pages 2 0 R >>\nendobj\n"
            b"2 0 obj\n<< /Type /Pages /Kids [3 0 R] /Count 1 /MediaBox [0 0 400 300] " + tree_res + b" >>\nendobj\n"
            b"3 0 obj\n<< /Type /Page /Parent 2 0 R /Contents 4 0 R " + page_res + b" >>\nendobj\n"
            b"4 0 obj\n<< /Length 5 0 R " + filt + b">>\nstream\n" + body + b"\nendstream\nendobj\n"
            b"5 0 obj\n" + str(len(body)).encode() + b"\nendobj\n" + extra_objects +
            b"trailer\n<< /Root 1 0 R /Size 6 >>\n%%EOF\n")


def graphic(content, **kw):
    return P.graphic_of_page(P.PdfFile(make_pdf(content, **kw)))[0]


def test_path_construction_follows_the_reference():
    g = graphic(b"10 20 m 110 20 l 110 120 l h f   5 5 m 6 6 l 7 5 l f*   1 2 3 4 re S")
    assert [e[0] for e in g] == ["Path"] * 3
    # f: `h` marks the subpath Closed without a closing segment, and f's own h leaves an empty Closed subpath (pdfgraphics.ml:932)
    _, w, sub, a = g[0]
    assert w == "NonZero" and sub == [("Closed", [("L", (10.0, 20.0), (110.0, 20.0)), ("L", (110.0, 20.0), (110.0, 120.0))]), ("Closed", [])]
    assert a["fill"] == ("/DeviceGray", [1.0]) and a["line"] is None and a["line_transparency"] == 1.0
    # f*: no h; the open segments become an Open subpath
    assert g[1][1] == "EvenOdd" and g[1][2] == [("Open", [("L", (5.0, 5.0), (6.0, 6.0)), ("L", (6.0, 6.0), (7.0, 5.0))])]
    # re = m l l l h: three segments (pdfgraphics.ml:1016-1025); S: stroke attributes only
    assert g[2][2] == [("Closed", [("L", (1.0, 2.0), (4.0, 2.0)), ("L", (4.0, 2.0), (4.0, 6.0)), ("L", (4.0, 6.0), (1.0, 6.0))])]
    assert g[2][3]["fill"] is None and g[2][3]["line"] == ("/DeviceGray", [1.0])


def test_curves_and_state():
    g = graphic(b"q 2 0 0 2 5 5 cm 0.2 0.4 0.6 rg 1 0 0 RG 3 w 1 j 2 J 4 M\n"
                b"0 0 m 1 1 2 2 3 3 c 4 4 5 5 v 6 6 7 7 y B Q 0 0 m 1 0 l 1 1 l b*")
    _, w, sub, a = g[0]
    assert sub == [("Open", [("C", (0.0, 0.0), (1.0, 1.0), (2.0, 2.0), (3.0, 3.0)), ("C", (3.0, 3.0), (3.0, 3.0), (4.0, 4.0), (5.0, 5.0)),
                             ("C", (5.0, 5.0), (6.0, 6.0), (7.0, 7.0), (7.0, 7.0))])]          # user space: cm moves nothing
    assert a["transform"] == (2.0, 0.0, 0.0, 2.0, 5.0, 5.0) and w == "NonZero"
    assert a["fill"] == ("/DeviceRGB", [0.2, 0.4, 0.6]) and a["line"] == ("/DeviceRGB", [1.0, 0.0, 0.0])
    assert (a["linewidth"], a["joinstyle"], a["capstyle"], a["mitrelimit"]) == (3.0, 1, 2, 4.0)
    # Q restores everything; b* = h B*, and B* closes once more
    _, w2, sub2, a2 = g[1]
    assert w2 == "EvenOdd" and [c for c, _ in sub2] == ["Closed", "Closed"] and sub2[1][1] == []
    assert a2["fill"] == ("/DeviceGray", [1.0]) and a2["linewidth"] == 1.0 and a2["transform"] == (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)


def test_reference_quirks_are_kept():
    # n without a pending clip leaves the partial path: it becomes an Open subpath of the next one (pdfgraphics.ml:1128-1129)
    g = graphic(b"0 0 m 9 9 l n 1 1 m 2 2 l 3 1 l f")
    assert g[0][2][0] == ("Open", [("L", (0.0, 0.0), (9.0, 9.0))]) and [c for c, _ in g[0][2]] == ["Open", "Closed"]
    # k stores c, y, m, k (pdfgraphics.ml:855); CS sets the NON-stroking space (826)
    g = graphic(b"0.1 0.2 0.3 0.4 k /DeviceRGB CS 0 0 m 1 1 l 2 0 l f")
    assert g[0][3]["fill"] == ("/DeviceRGB", [0.1, 0.3, 0.2, 0.4])
    # ExtGState: /LC sets cap and join, /LJ is not read, /CA and /ca count only as reals (pdfgraphics.ml:522-556)
    g = graphic(b"/G0 gs /G1 gs 0 0 m 1 1 l B", extgstate="/G0 << /LC 2 /LJ 1 /CA 1 /ca 0.5 /LW 7 /ML 3.5 >> /G1 << /CA 0.25 >>")
    a = g[0][3]
    assert (a["capstyle"], a["joinstyle"], a["linewidth"], a["mitrelimit"]) == (2, 2, 7.0, 3.5)
    assert (a["line_transparency"], a["fill_transparency"]) == (0.25, 0.5)
    with pytest.raises(P.PdfError):
        graphic(b"1 1 l f")                     # Pdfgraphics: Op_l outside a path
    with pytest.raises(P.PdfError):
        graphic(b"Q")                           # Unbalanced q/Q Ops


def test_clips_marked_content_forms_and_text():
    form = b"7 0 obj\n<< /Type /XObject /Subtype /Form /BBox [0 0 9 9] /Matrix [9 0 0 9 0 0] /Length 34 >>\nstream\n0 1 0 rg 0 0 m 5 0 l 5 5 l h f 2 w\nendstream\nendobj\n"
    g = graphic(b"/L BDC q 0 0 m 10 0 l 10 10 l h W* n 1 1 m 2 2 l 3 1 l f /Fm1 Do Q 4 4 m 5 5 l 6 4 l S EMC "
                b"BT 0 0 1 rg (x) Tj ET 0 0 m 1 1 l 2 0 l f", xobjects="/Fm1 7 0 R", extra_objects=form)
    assert [e[0] for e in g] == ["MCSection", "Text", "Path"]
    mc = g[0][1]
    assert [e[0] for e in mc] == ["Clip", "Path"]
    clip = mc[0]
    assert clip[1][0] == "EvenOdd" and [e[0] for e in clip[2]] == ["Path", "Path"]     # everything up to the matching Q
    assert clip[2][1][3]["fill"] == ("/DeviceRGB", [0.0, 1.0, 0.0])                     # the form's content, its /Matrix not read
    assert mc[1][3]["linewidth"] == 1.0                                                 # Q put back what the form set (2 w)
    assert g[2][3]["fill"] == ("/DeviceRGB", [0.0, 0.0, 1.0])                           # colour set inside BT ... ET stays
    objs = P.scene_of_graphic(g)
    assert [o[0] for o in objs] == ["fill", "fill", "stroke", "fill"]                   # clips and sections descended into, text dropped


def test_scene_of_graphic_order_and_colours():
    g = graphic(b"/G0 gs 1 0 0 rg 0 0 1 RG 0 0 m 50 0 l 50 50 l h B  /Cs1 cs 0.5 0.25 0.125 sc 0 0 m 9 0 l 9 9 l f  /Cs2 cs 0.5 sc 0 0 m 9 0 l 9 9 l f",
                extgstate="/G0 << /ca 0.5 /CA 0.75 >>", colorspaces="/Cs1 [/ICCBased 8 0 R] /Cs2 [/Separation /Spot /DeviceGray 9 0 R]",
                extra_objects=b"8 0 obj\n<< /N 3 /Alternate /DeviceRGB /Length 0 >>\nstream\n\nendstream\nendobj\n")
    objs = P.scene_of_graphic(g)
    assert [o[0] for o in objs] == ["stroke", "fill", "fill", "fill"]                   # line @ fill (render.ml:1553)
    want_line = S.dissolve(S.colour_of_rgba_float(0.0, 0.0, 1.0, 1.0), int(0.75 * 255.0))
    want_fill = S.dissolve(S.colour_of_rgba_float(1.0, 0.0, 0.0, 1.0), int(0.5 * 255.0))
    assert objs[0][1].c0 == want_line and objs[1][1].c0 == want_fill
    assert objs[0][4] == (abi.CAP_BUTT, abi.JOIN_MITRED, abi.CAP_BUTT, 10.0, 1.0)       # render.ml:1511-1522
    assert objs[2][1].c0 == S.dissolve(S.colour_of_rgba_float(0.5, 0.25, 0.125, 1.0), int(0.5 * 255.0))   # ICCBased -> its alternate
    assert objs[3][1].c0 == P.RED                                                        # not handled: red (render.ml:1513)


def test_reader_details():
    # uncompressed stream, resources on the page itself, names with #xx, strings with escapes, comments
    pdf = P.PdfFile(make_pdf(b"% a comment\n/G#30 gs (a\\)b\\051) Tj <41 4> Tj 0 0 m 1 1 l 2 0 l f", extgstate="/G0 << /ca 0.5 >>", compress=False, inherit=False))
    g, box = P.graphic_of_page(pdf)
    assert box == [0.0, 0.0, 400.0, 300.0] and g[0][3]["fill_transparency"] == 0.5
    ops = pdf.operators(pdf.first_page()[1])
    assert ops[1] == ("Tj", [b"a)b)"]) and ops[2] == ("Tj", [b"A@"])
    with pytest.raises(P.PdfError):
        P.PdfFile(b"%PDF-1.5 nothing here")
    bad = make_pdf(b"0 0 m f").replace(b"/FlateDecode", b"/LZWDecode  ")
    with pytest.raises(P.PdfError):
        P.graphic_of_page(P.PdfFile(bad))


def _page_scene(ctx=None):
    content = (b"/G0 gs 0.9 0.9 0.2 rg 20 20 m 380 30 l 370 280 l 30 270 l 20 20 l f\n"
               b"q /G1 gs 0.1 0.3 0.8 rg 0.8 0.1 0.1 RG 9 w 1 j 1 J 60 60 m 150 260 200 40 340 240 c 340 240 m 300 70 l B Q\n"
               b"0 0 0 rg 100 100 m 300 100 l 300 200 l 100 200 l 100 100 l 150 130 m 250 130 l 250 170 l 150 170 l 150 130 l f*\n"
               b"0 0.5 0 RG 4 w 2 J 0 j 40 150 m 120 150 l 120 90 l S /Fm1 Do")
    form = b"7 0 obj\n<< /Subtype /Form /Length 60 >>\nstream\n0.5 g 200 210 m 260 210 l 260 260 l 200 260 l 200 210 l f      \nendstream\nendobj\n"
    pdf = P.PdfFile(make_pdf(content, extgstate="/G0 << /ca 1.0 >> /G1 << /ca 0.6 /CA 0.7 >>", xobjects="/Fm1 7 0 R", extra_objects=form))
    b = S.SceneBuilder()
    n = P.add_pdf_page(b, pdf, scale=0.8, origin=(3.25, 2.5), flip_height=300.0, ctx=ctx)
    assert n == 6
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, 330.0, 250.0)
    return b


def test_stroke_edges_on_the_host_equal_the_oracle(oracle):
    spec = abi.strokespec(abi.CAP_ROUND, abi.JOIN_ROUND, abi.CAP_ROUND, 10.0, 7.2)
    path = [[("C", (51.25, 194.5), (123.25, 34.5), (163.25, 210.5), (275.25, 50.5))], [("L", (275.25, 50.5), (243.25, 186.5))]]
    rec, cnt = abi._path_records(path)
    _, _, _, ref_edges = oracle.strokepath((spec.startcap, spec.join, spec.endcap, spec.mitrelimit, spec.linewidth), rec, cnt)
    assert np.array_equal(P.stroke_edges_host(spec, path), np.asarray(ref_edges, dtype=np.int32).reshape(-1, 4))
    objs, n, nbg, edges, points = _page_scene().arrays()       # the whole page builds without a device
    assert n - nbg >= 7 and len(edges) > 100


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's fixture PDFs exist in the build container only")
def test_reference_fixture_pdfs_render_through_the_oracle(oracle):
    """Every fixture PDF of the reference: importer -> stroker -> flattening -> the oracle's render_frame (the icons are
    stroked and filled paths, the text pages glyph outlines); something other than the background comes out."""
    for f in sorted(glob.glob(os.path.join(REF, "*.pdf"))):
        pdf = P.PdfFile.open(f)
        elts, box = P.graphic_of_page(pdf)
        pts = [p for e in P.scene_of_graphic(elts) for sp in e[3] for s in sp for p in s[1:]]
        xs, ys = [p[0] for p in pts], [p[1] for p in pts]
        sc = min(300.0 / max(max(xs) - min(xs), 1e-9), 200.0 / max(max(ys) - min(ys), 1e-9))
        W, H = int((max(xs) - min(xs)) * sc) + 20, int((max(ys) - min(ys)) * sc) + 20
        b = S.SceneBuilder()
        P.add_pdf_page(b, pdf, scale=sc, origin=(10 - sc * min(xs), 10 - sc * (box[3] - max(ys))), flip_height=box[3])
        b.begin_background()
        b.rectangle(S.rgba8(200, 220, 240), 0.0, 0.0, float(W), float(H))
        objs, n, nbg, edges, points = b.arrays()
        ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
        assert (ref != ref[0, 0]).sum() > 50, os.path.basename(f)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's fixture PDFs exist in the build container only")
def test_reference_fixture_pdfs():
    import json

    # lion.pdf -> exactly the committed fixture (coherence_renderer_b200/scenes/lion.json, tools/make_lion_fixture.py)
    objs = P.scene_of_graphic(P.graphic_of_page(P.PdfFile.open(os.path.join(REF, "lion.pdf")))[0])
    with open(os.path.join(os.path.dirname(S.__file__), "scenes", "lion.json")) as f:
        lion = json.load(f)["paths"]
    assert len(objs) == len(lion) == 132
    for (kind, fill, winding, subpaths, spec), want in zip(objs, lion):
        assert kind == "fill" and winding == "NonZero"
        assert fill.c0 == S.colour_of_rgba_float(*want["rgb"], 1.0)
        pts = [list(subpaths[0][0][1])] + [list(s[2]) for s in subpaths[0]]
        assert pts == want["subpaths"][0] and all(not sp for sp in subpaths[1:])
    # every fixture imports; painted paths = painting operators in its streams
    counts = {}
    for f in sorted(glob.glob(os.path.join(REF, "*.pdf"))):
        objs = P.scene_of_graphic(P.graphic_of_page(P.PdfFile.open(f))[0])
        counts[os.path.basename(f)] = (sum(o[0] == "fill" for o in objs), sum(o[0] == "stroke" for o in objs))
    # (fills, strokes) = what the painting operators of each file ask for: f / f* one fill, S one stroke, B* both
    assert counts == {"aatext.pdf": (51, 0), "brushcurve.pdf": (0, 1), "down.pdf": (2, 1), "filtertext1.pdf": (86, 0), "filtertext2.pdf": (106, 0),
                      "lion.pdf": (132, 0), "lionfilter1.pdf": (99, 0), "lionfilter2.pdf": (44, 0), "logo.pdf": (1, 0), "mintext1.pdf": (131, 0),
                      "mintext2.pdf": (71, 0), "pointer.pdf": (1, 1), "q.pdf": (1, 0), "up.pdf": (2, 1), "zoom.pdf": (1, 2)}
    # a text page builds into a scene on the host (the stroker and the flattening of N2 included for the icons)
    for name in ("mintext1.pdf", "up.pdf", "zoom.pdf"):
        b = S.SceneBuilder()
        P.add_pdf_page(b, P.PdfFile.open(os.path.join(REF, name)), scale=1.5, flip_height=800.0)
        objs_, n, nbg, edges, _ = b.arrays()
        assert n >= 3 and len(edges) > 10


@pytest.mark.gpu
def test_imported_page_against_oracle(ctx, oracle):
    W, H = 330, 250
    bh = _page_scene()
    bd = _page_scene(ctx)                                       # stroke outlines flattened and sorted on the device
    objs, n, nbg, edges, points = bd.arrays()
    assert np.array_equal(edges, bh.arrays()[3])                # host route = device route
    ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    got = ctx.fb_read_rgba(0, 0, W, H)
    assert np.array_equal(np.asarray(got).view(np.uint32).reshape(H, W), ref)
    assert len(np.unique(ref)) > 40                             # fills, strokes, translucency, antialiased edges
    ctx.scene_free(sc)


@pytest.mark.gpu
@pytest.mark.parametrize("scale,W,H", [(2.2, 900, 110), (0.55, 240, 40), (7.0, 1024, 96)])
def test_text_page_against_oracle(ctx, oracle, scale, W, H):
    """The reference's text fixture (mintext1.pdf through pdf_import, committed as scenes/mintext1.json): glyph outlines at
    reading size, as tiny text (many crossings in one 32-pixel window) and enlarged (long curves), GPU = oracle."""
    b = S.SceneBuilder()
    S.add_text_page(b, scale, origin=(-scale * 20.0, -scale * 60.0))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, edges, points = b.arrays()
    ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    got = np.asarray(ctx.fb_read_rgba(0, 0, W, H)).view(np.uint32).reshape(H, W)
    assert np.array_equal(got, ref)
    assert (ref != ref[0, 0]).sum() > W * H // 40              # there is text in the frame
    ctx.scene_free(sc)
