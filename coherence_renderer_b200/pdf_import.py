"""N4 — PDF import: the scene of a PDF page, host side, in front of the raster path.

Three layers, as in the reference:

* ``PdfFile`` — a MINIMAL reader of PDF files (objects, FlateDecode streams, first page of the page tree, inherited
  resources, content-stream operators).  In the reference this is the third-party library camlpdf (``Pdfread``,
  ``Pdfpage``, ``Pdfops``; not under /root/reference, unpinned in its Makefile); it is not restated, only stood in for:
  files with object streams, encryption or filters other than FlateDecode are refused loudly.
* ``graphic_of_page`` — ``Pdfgraphics.graphic_of_page`` (pdfgraphics.ml:727-1245), the path side of it: the graphics state
  with its q / Q stack, path construction and painting, marked-content sections, clips (kept as a structure, not applied),
  form XObjects.  The reference's behaviour is kept where it differs from the PDF specification, because the scenes the
  raster path renders are the ones THIS reader produces:
    - coordinates stay in user space; ``cm`` only updates ``path_transform``, which Render.scene_of_graphic ignores;
    - ``h`` marks the subpath Closed but adds no closing segment, and Polygon.edgelist_of_path reads segments only;
    - ``f`` runs ``h`` first (pdfgraphics.ml:932), leaving an empty Closed subpath behind a path that ended in ``h``;
    - ``n`` without a pending clip leaves the partial path in place (pdfgraphics.ml:1128-1129);
    - ``K`` / ``k`` store c, y, m, k (pdfgraphics.ml:851, 855); ``CS`` sets the NON-stroking colour space (826);
    - an ExtGState's /LC sets cap AND join, /LJ is not read (pdfgraphics.ml:547-556); /CA and /ca count only as reals.
* ``scene_of_graphic`` — ``Render.scene_of_graphic`` (render.ml:1476-1565): per painted path a StrokedPath object (the stroker
  of N2) listed BEFORE the filled Path object (``line @ fill``, render.ml:1553 — after load_text's ``rev`` the fill ends up in
  front of its own stroke, which is the reference's order, not PDF's); MCSections and Clips are descended into; everything
  else is dropped.  ``add_pdf_page`` appends ``Group (rev objs)`` (examples.ml:158-165, 174-180) to a SceneBuilder.

Text objects, images, shadings and patterns produce nothing in Render.scene_of_graphic and are skipped here as well.
"""
import re
import zlib

import numpy as np

from . import abi
from . import scene as S

__all__ = ["PdfFile", "PdfError", "graphic_of_page", "scene_of_graphic", "add_pdf_page", "stroke_edges_host"]


class PdfError(Exception):
    pass


class Name(str):
    pass


class Ref(tuple):
    pass


class Stream:
    def __init__(self, d, raw):
        self.dict, self.raw = d, raw


_WS = b"\x00\t\n\x0c\r "
_DELIM = b"()<>[]{}/%"
_NUM = re.compile(rb"^[+-]?(\d+\.?\d*|\.\d+)$")


class _Lexer:
    """Tokens of PDF syntax (objects and content streams)."""

    def __init__(self, data, pos=0):
        self.d, self.p = data, pos

    def skip_ws(self):
        d, n = self.d, len(self.d)
        while self.p < n:
            c = d[self.p]
            if c in _WS:
                self.p += 1
            elif c == 0x25:  # % comment
                while self.p < n and d[self.p] not in b"\r\n":
                    self.p += 1
            else:
                break

    def token(self):
        """None at the end; otherwise ('num', int | float), ('name', Name), ('str', bytes), ('kw', bytes) or a delimiter."""
        self.skip_ws()
        d, n = self.d, len(self.d)
        if self.p >= n:
            return None
        c = d[self.p : self.p + 1]
        if c == b"/":
            q = self.p + 1
            while q < n and d[q] not in _WS and d[q] not in _DELIM:
                q += 1
            raw = d[self.p + 1 : q]
            self.p = q
            raw = re.sub(rb"#([0-9A-Fa-f]{2})", lambda m: bytes([int(m.group(1), 16)]), raw)
            return ("name", Name("/" + raw.decode("latin-1")))
        if c == b"(":
            depth, q, out = 1, self.p + 1, bytearray()
            while q < n and depth:
                ch = d[q]
                if ch == 0x5C:  # backslash
                    q += 1
                    e = d[q : q + 1]
                    if e in b"nrtbf":
                        out += {b"n": b"\n", b"r": b"\r", b"t": b"\t", b"b": b"\b", b"f": b"\f"}[e]
                    elif e.isdigit():
                        m = re.match(rb"[0-7]{1,3}", d[q : q + 3])
                        out.append(int(m.group(0), 8) & 255)
                        q += len(m.group(0)) - 1
                    elif e in b"\r\n":
                        if e == b"\r" and d[q + 1 : q + 2] == b"\n":
                            q += 1
                    else:
                        out += e
                elif ch == 0x28:
                    depth += 1
                    out.append(ch)
                elif ch == 0x29:
                    depth -= 1
                    if depth:
                        out.append(ch)
                else:
                    out.append(ch)
                q += 1
            self.p = q
            return ("str", bytes(out))
        if c == b"<":
            if d[self.p : self.p + 2] == b"<<":
                self.p += 2
                return ("<<",)
            q = d.index(b">", self.p)
            hx = re.sub(rb"\s", b"", d[self.p + 1 : q])
            self.p = q + 1
            return ("str", bytes.fromhex((hx + b"0" * (len(hx) & 1)).decode("ascii")))
        if c == b">":
            if d[self.p : self.p + 2] != b">>":
                raise PdfError("stray '>'")
            self.p += 2
            return (">>",)
        if c in (b"[", b"]", b"{", b"}"):
            self.p += 1
            return (c.decode(),)
        q = self.p
        while q < n and d[q] not in _WS and d[q] not in _DELIM:
            q += 1
        w = d[self.p : q]
        if not w:
            raise PdfError("unexpected byte %r" % c)
        self.p = q
        if _NUM.match(w):
            return ("num", float(w) if b"." in w else int(w))
        return ("kw", w)


def _parse_object(lx):
    """One PDF object at the lexer's position (indirect references resolved lazily as Ref)."""
    t = lx.token()
    if t is None:
        raise PdfError("object expected")
    k = t[0]
    if k == "num":
        if isinstance(t[1], int) and t[1] >= 0:  # maybe `n g R`
            save = lx.p
            t2 = lx.token()
            if t2 and t2[0] == "num" and isinstance(t2[1], int):
                t3 = lx.token()
                if t3 == ("kw", b"R"):
                    return Ref((t[1], t2[1]))
            lx.p = save
        return t[1]
    if k in ("name", "str"):
        return t[1]
    if k == "kw":
        if t[1] in (b"true", b"false"):
            return t[1] == b"true"
        if t[1] == b"null":
            return None
        raise PdfError("unexpected keyword %r" % t[1])
    if k == "[":
        out = []
        while True:
            lx.skip_ws()
            if lx.d[lx.p : lx.p + 1] == b"]":
                lx.p += 1
                return out
            out.append(_parse_object(lx))
    if k == "<<":
        out = {}
        while True:
            lx.skip_ws()
            if lx.d[lx.p : lx.p + 2] == b">>":
                lx.p += 2
                return out
            key = lx.token()
            if key[0] != "name":
                raise PdfError("dictionary key expected")
            out[key[1]] = _parse_object(lx)
    raise PdfError("unexpected token %r" % (t,))


class PdfFile:
    """Stand-in for camlpdf's Pdfread / Pdfpage on simple files (classic cross-reference tables are not even needed: the
    objects are found by their `n g obj` headers, the last definition of a number winning as with incremental updates)."""

    def __init__(self, data):
        self.data = data
        if b"/Encrypt" in data[data.rfind(b"trailer") :] if b"trailer" in data else False:
            raise PdfError("Pdfgraphics: File is encrypted")
        self._where = {}
        for m in re.finditer(rb"(?<![0-9])(\d+)[ \t\r\n]+(\d+)[ \t\r\n]+obj\b", data):
            self._where[int(m.group(1))] = m.end()
        self._cache = {}
        if not self._where:
            raise PdfError("no objects found (object streams are not supported by this reader)")

    @classmethod
    def open(cls, path):
        with open(path, "rb") as f:
            return cls(f.read())

    def obj(self, n):
        if n in self._cache:
            return self._cache[n]
        if n not in self._where:
            return None
        lx = _Lexer(self.data, self._where[n])
        o = _parse_object(lx)
        if isinstance(o, dict):
            lx.skip_ws()
            if self.data[lx.p : lx.p + 6] == b"stream":
                p = lx.p + 6
                if self.data[p : p + 2] == b"\r\n":
                    p += 2
                elif self.data[p : p + 1] in (b"\n", b"\r"):
                    p += 1
                length = self.direct(o.get("/Length"))
                if isinstance(length, int) and self.data[p + length : p + length + 20].lstrip(_WS).startswith(b"endstream"):
                    raw = self.data[p : p + length]
                else:
                    e = self.data.index(b"endstream", p)
                    raw = self.data[p:e].rstrip(b"\r\n")
                o = Stream(o, raw)
        self._cache[n] = o
        return o

    def direct(self, o):
        seen = 0
        while isinstance(o, Ref):
            o = self.obj(o[0])
            seen += 1
            if seen > 64:
                raise PdfError("reference loop")
        return o

    def lookup(self, d, key):
        d = self.direct(d)
        if isinstance(d, Stream):
            d = d.dict
        if not isinstance(d, dict):
            return None
        return self.direct(d.get(key))

    def stream_data(self, s):
        s = self.direct(s)
        if not isinstance(s, Stream):
            raise PdfError("stream expected")
        f = self.direct(s.dict.get("/Filter"))
        filters = [] if f is None else ([self.direct(x) for x in f] if isinstance(f, list) else [f])
        data = s.raw
        for name in filters:
            if name != "/FlateDecode":
                raise PdfError("filter %s is not supported by this reader" % name)
            parms = self.lookup(s.dict, "/DecodeParms")
            if isinstance(parms, dict) and self.direct(parms.get("/Predictor", 1)) not in (1, None):
                raise PdfError("predictors are not supported by this reader")
            data = zlib.decompressobj().decompress(data)
        return data

    def trailer_root(self):
        roots = re.findall(rb"/Root[ \t\r\n]+(\d+)[ \t\r\n]+(\d+)[ \t\r\n]+R", self.data)
        if not roots:
            raise PdfError("no /Root")
        return self.obj(int(roots[-1][0]))

    def first_page(self):
        """(resources, [content streams], mediabox) of the first page (Pdfpage.pages_of_pagetree: /Resources and /MediaBox
        are inherited down the tree)."""
        node = self.lookup(self.trailer_root(), "/Pages")
        res, box = None, None
        for _ in range(64):
            if not isinstance(node, dict):
                raise PdfError("bad page tree")
            res = self.lookup(node, "/Resources") if "/Resources" in node else res
            box = self.lookup(node, "/MediaBox") if "/MediaBox" in node else box
            if node.get("/Type") == "/Page" or "/Kids" not in node:
                c = node.get("/Contents")
                cd = self.direct(c)
                contents = [] if cd is None else (list(cd) if isinstance(cd, list) else [c])
                return (res if isinstance(res, dict) else {}), contents, [float(self.direct(v)) for v in (box or [0, 0, 612, 792])]
            kids = self.lookup(node, "/Kids")
            if not kids:
                raise PdfError("No pages in PDF file")
            node = self.direct(kids[0])
        raise PdfError("page tree too deep")

    def operators(self, contents):
        """Pdfops.parse_operators on the concatenated content streams: [(operator, [operands])]."""
        data = b"\n".join(self.stream_data(c) for c in contents)
        lx, ops, stack = _Lexer(data), [], []
        while True:
            save = lx.p
            t = lx.token()
            if t is None:
                return ops
            if t[0] == "kw" and t[1] not in (b"true", b"false", b"null"):
                op = t[1].decode("latin-1")
                if op == "BI":  # inline image: skipped up to EI
                    m = re.search(rb"[\s]EI(?=[\s]|$)", data[lx.p :])
                    if not m:
                        raise PdfError("inline image without EI")
                    lx.p += m.end()
                    ops.append(("InlineImage", []))
                else:
                    ops.append((op, stack))
                stack = []
            else:
                lx.p = save
                stack.append(_parse_object(lx))


# ---------------------------------------------------------------- Pdfgraphics.graphic_of_page (path side)
def _default_state():  # pdfgraphics.ml:315-345
    return dict(objectclass="page", clip=None, fill=[1.0], line=[1.0], linewidth=1.0, mitrelimit=10.0, joinstyle=0, capstyle=0,
                cs_stroke="/DeviceGray", cs_nonstroke="/DeviceGray", dash=([], 0.0), transform=(1.0, 0.0, 0.0, 1.0, 0.0, 0.0),
                opacity_stroke=1.0, opacity_nonstroke=1.0)


def _compose(m, t):  # Pdftransform.matrix_compose m t: t applied first (a point is transformed by t, then by m)
    a, b, c, d, e, f = m
    a2, b2, c2, d2, e2, f2 = t
    return (a * a2 + c * b2, b * a2 + d * b2, a * c2 + c * d2, b * c2 + d * d2, a * e2 + c * f2 + e, b * e2 + d * f2 + f)


class _Interp:
    def __init__(self, pdf, resources):
        self.pdf, self.state, self.stack, self.depth = pdf, _default_state(), [], 0

    # -- Pdfspace.read_colourspace, as far as Render.fill_of_pdf_colour distinguishes spaces (render.ml:1481-1513)
    def colourspace(self, resources, name):
        if name in ("/DeviceGray", "/DeviceRGB", "/DeviceCMYK"):
            return str(name)
        cs = self.pdf.lookup(self.pdf.lookup(resources, "/ColorSpace"), name)
        return self._space_of_object(cs)

    def _space_of_object(self, cs):
        cs = self.pdf.direct(cs)
        if isinstance(cs, Name):
            return str(cs) if cs in ("/DeviceGray", "/DeviceRGB", "/DeviceCMYK") else "other"
        if isinstance(cs, list) and cs and self.pdf.direct(cs[0]) == "/ICCBased" and len(cs) > 1:
            st = self.pdf.direct(cs[1])
            alt = self.pdf.lookup(st, "/Alternate")
            if alt is not None:
                return self._space_of_object(alt)  # render.ml:1510-1511: the alternate decides
            return {1: "/DeviceGray", 3: "/DeviceRGB", 4: "/DeviceCMYK"}.get(self.pdf.lookup(st, "/N"), "other")
        return "other"

    def attrs(self, fill, stroke):  # pdfgraphics.ml:356-396
        s = self.state
        return dict(transform=s["transform"], fill=(s["cs_nonstroke"], list(s["fill"])) if fill else None,
                    line=(s["cs_stroke"], list(s["line"])) if stroke else None, linewidth=s["linewidth"], joinstyle=s["joinstyle"],
                    capstyle=s["capstyle"], dash=s["dash"], mitrelimit=s["mitrelimit"],
                    fill_transparency=s["opacity_nonstroke"] if fill else 1.0, line_transparency=s["opacity_stroke"] if stroke else 1.0)

    def gs(self, resources, name):  # pdfgraphics.ml:517-576
        g = self.pdf.lookup(self.pdf.lookup(resources, "/ExtGState"), name)
        if not isinstance(g, dict):
            raise PdfError("Bad Op_gs")
        s, look = self.state, lambda k: self.pdf.direct(g.get(k))
        if isinstance(look("/CA"), float):
            s["opacity_stroke"] = look("/CA")
        if isinstance(look("/ca"), float):
            s["opacity_nonstroke"] = look("/ca")
        lw = look("/LW")
        if isinstance(lw, (int, float)) and not isinstance(lw, bool):
            s["linewidth"] = float(lw)
        lc = look("/LC")
        if isinstance(lc, int) and not isinstance(lc, bool):
            s["capstyle"] = lc
            s["joinstyle"] = lc   # (the reference reads /LC twice)
        ml = look("/ML")
        if isinstance(ml, (int, float)) and not isinstance(ml, bool):
            s["mitrelimit"] = float(ml)

    def process_ops(self, resources, partial, ops):
        """pdfgraphics.ml:1120-1188; returns (partial, elements in paint order)."""
        graphic, i, n = [], 0, len(ops)
        while i < n:
            op, args = ops[i]
            i += 1
            if op == "n":
                clip = self.state["clip"]
                if clip is None:
                    continue   # the partial path stays where it is
                level, j = 0, i
                while j < n:   # getuntil_matching_Q
                    if ops[j][0] == "q":
                        level += 1
                    elif ops[j][0] == "Q":
                        if level == 0:
                            break
                        level -= 1
                    j += 1
                _, elts = self.process_ops(resources, None, ops[i:j])
                graphic.append(("Clip", clip, elts))
                partial, i = None, j
            elif op in ("BMC", "BDC"):
                level, j, found = 0, i, False
                while j < n:   # getuntil_matching_emc
                    if ops[j][0] in ("BMC", "BDC"):
                        level += 1
                    elif ops[j][0] == "EMC":
                        if level == 0:
                            found = True
                            break
                        level -= 1
                    j += 1
                if not found:
                    continue   # Missing EMC: the operator alone is dropped
                partial, elts = self.process_ops(resources, partial, ops[i:j])
                graphic.append(("MCSection", elts))
                i = j + 1
            elif op == "BT":
                j = i
                while j < n and ops[j][0] != "ET":
                    j += 1
                self.state["objectclass"] = "text"
                self.process_ops(resources, "text", ops[i:j])   # state changes inside have global effect
                graphic.append(("Text",))
                partial, i = "text", j
            elif op == "ET":
                self.state["objectclass"] = "page"
            else:
                partial = self.process_op(resources, partial, graphic, op, args)
        return partial, graphic

    def _need_path(self, partial, op):
        if self.state["objectclass"] != "path" or not isinstance(partial, list):
            raise PdfError("Pdfgraphics: Op_" + op)

    def process_op(self, resources, partial, graphic, op, a):
        """pdfgraphics.ml:727-1097.  partial: None | "text" | [sp, cp, segs (in order), subpaths (in order)]."""
        s = self.state
        fl = lambda k: float(a[k])
        if op in ("W", "W*"):
            if isinstance(partial, list) and (partial[2] or partial[3]):
                path = partial[3] + ([("Closed", list(partial[2]))] if partial[2] else [])
                s["clip"] = ("NonZero" if op == "W" else "EvenOdd", path)
        elif op == "j":
            s["joinstyle"] = int(a[0])
        elif op == "J":
            s["capstyle"] = int(a[0])
        elif op == "w":
            s["linewidth"] = fl(0)
        elif op == "M":
            s["mitrelimit"] = fl(0)
        elif op == "q":
            c = dict(s)
            self.stack.append(c)
        elif op == "Q":
            if not self.stack:
                raise PdfError("Unbalanced q/Q Ops")
            self.state = self.stack.pop()
        elif op in ("SC", "SCN"):
            if a and isinstance(a[-1], Name):
                pass   # a pattern or named colour: nothing Render.fill_of_pdf_colour can use
            else:
                s["line"] = [float(v) for v in a]
        elif op in ("sc", "scn"):
            if a and isinstance(a[-1], Name):
                s["fill"] = "named"
            else:
                s["fill"] = [float(v) for v in a]
        elif op in ("CS", "cs"):
            s["cs_nonstroke"] = self.colourspace(resources, a[0])   # CS too (pdfgraphics.ml:824-827)
        elif op == "G":
            s["cs_stroke"], s["line"] = "/DeviceGray", [fl(0)]
        elif op == "g":
            s["cs_nonstroke"], s["fill"] = "/DeviceGray", [fl(0)]
        elif op == "RG":
            s["cs_stroke"], s["line"] = "/DeviceRGB", [fl(0), fl(1), fl(2)]
        elif op == "rg":
            s["cs_nonstroke"], s["fill"] = "/DeviceRGB", [fl(0), fl(1), fl(2)]
        elif op == "K":
            s["cs_stroke"], s["line"] = "/DeviceCMYK", [fl(0), fl(2), fl(1), fl(3)]
        elif op == "k":
            s["cs_nonstroke"], s["fill"] = "/DeviceCMYK", [fl(0), fl(2), fl(1), fl(3)]
        elif op == "gs":
            self.gs(resources, a[0])
        elif op == "m":
            s["objectclass"] = "path"
            p = (fl(0), fl(1))
            if isinstance(partial, list):
                subpaths = partial[3] + ([("Open", list(partial[2]))] if partial[2] else [])
                return [p, p, [], subpaths]
            return [p, p, [], []]
        elif op == "l":
            self._need_path(partial, op)
            p = (fl(0), fl(1))
            return [partial[0], p, partial[2] + [("L", partial[1], p)], partial[3]]
        elif op == "c":
            self._need_path(partial, op)
            ep = (fl(4), fl(5))
            return [partial[0], ep, partial[2] + [("C", partial[1], (fl(0), fl(1)), (fl(2), fl(3)), ep)], partial[3]]
        elif op == "v":
            self._need_path(partial, op)
            ep = (fl(2), fl(3))
            return [partial[0], ep, partial[2] + [("C", partial[1], partial[1], (fl(0), fl(1)), ep)], partial[3]]
        elif op == "y":
            self._need_path(partial, op)
            ep = (fl(2), fl(3))
            return [partial[0], ep, partial[2] + [("C", partial[1], (fl(0), fl(1)), ep, ep)], partial[3]]
        elif op == "h":
            self._need_path(partial, op)
            return [partial[0], partial[1], [], partial[3] + [("Closed", list(partial[2]))]]
        elif op in ("s", "b", "b*"):
            partial = self.process_op(resources, partial, graphic, "h", [])
            return self.process_op(resources, partial, graphic, {"s": "S", "b": "B", "b*": "B*"}[op], [])
        elif op in ("f", "F", "B*"):   # these close the current subpath first
            if s["objectclass"] != "path":
                raise PdfError("Pdfgraphics: Op_" + op)
            partial = self.process_op(resources, partial, graphic, "h", [])
            s["objectclass"] = "page"
            subpaths = partial[3] + ([("Open", list(partial[2]))] if partial[2] else [])
            if op == "B*":
                graphic.append(("Path", "EvenOdd", subpaths, self.attrs(True, True)))
            else:
                graphic.append(("Path", "NonZero", partial[3], self.attrs(True, False)))
            return [partial[0], partial[1], [], []]
        elif op in ("f*", "S", "B"):
            self._need_path(partial, op)
            s["objectclass"] = "page"
            subpaths = partial[3] + ([("Open", list(partial[2]))] if partial[2] else [])
            winding, at = {"f*": ("EvenOdd", (True, False)), "S": ("EvenOdd", (False, True)), "B": ("NonZero", (True, True))}[op]
            graphic.append(("Path", winding, subpaths, self.attrs(*at)))
            return [partial[0], partial[1], [], []]
        elif op == "re":
            x, y, w, h = fl(0), fl(1), fl(2), fl(3)
            for o2, a2 in (("m", [x, y]), ("l", [x + w, y]), ("l", [x + w, y + h]), ("l", [x, y + h]), ("h", [])):
                partial = self.process_op(resources, partial, graphic, o2, a2)
            return partial
        elif op == "Do":
            x = self.pdf.lookup(self.pdf.lookup(resources, "/XObject"), a[0])
            if not isinstance(x, Stream):
                raise PdfError("Unknown xobject")
            sub = x.dict.get("/Subtype")
            if sub == "/Image":
                graphic.append(("Image",))
            elif sub == "/Form":
                # read_form_xobject (pdfgraphics.ml:1197-1226): the form's resources over the page's; its /Matrix is not
                # read and no state is pushed: what the form sets stays set
                if self.depth > 32:
                    raise PdfError("form XObjects nested too deeply")
                merged = dict(resources)
                own = self.pdf.lookup(x.dict, "/Resources")
                if isinstance(own, dict):
                    merged.update(own)
                self.depth += 1
                _, elts = self.process_ops(merged, None, self.pdf.operators([x]))
                self.depth -= 1
                graphic.extend(elts)
            else:
                raise PdfError("Unknown kind of xobject")
        elif op == "cm":
            s["transform"] = _compose(s["transform"], tuple(float(v) for v in a))
        elif op == "d":
            s["dash"] = ([float(v) for v in a[0]], float(a[1]))
        elif op == "InlineImage":
            graphic.append(("Image",))
            return None
        elif op in ("MP", "DP"):
            if s["objectclass"] == "page":
                graphic.append(("MCPoint",))
                return None
        elif op == "sh":
            graphic.append(("Shading",))
        # everything else (text state, ri, i, BX / EX, unknown operators) leaves paths alone
        return partial


def graphic_of_page(pdf):
    """Pdfgraphics.graphic_of_page of the file's first page: (elements, mediabox)."""
    resources, contents, box = pdf.first_page()
    it = _Interp(pdf, resources)
    _, elts = it.process_ops(resources, None, pdf.operators(contents))
    return elts, box


# ---------------------------------------------------------------- Render.scene_of_graphic
def _rgb_of_cmyk(c, m, y, k):  # render.ml:1476-1479
    return 1.0 - min(1.0, c * (1.0 - k) + k), 1.0 - min(1.0, m * (1.0 - k) + k), 1.0 - min(1.0, y * (1.0 - k) + k)


RED = S.rgba8(255, 0, 0)


def fill_of_pdf_colour(space, vals, transparency):
    """render.ml:1481-1513: the plain fill of a PDF colour dissolved by the path's transparency; red for what is not handled."""
    delta = int(transparency * 255.0)
    if not isinstance(vals, list):
        return S.Fill.plain(RED)
    if space == "/DeviceRGB" and len(vals) == 3:
        r, g, b = vals
    elif space == "/DeviceCMYK" and len(vals) == 4:
        r, g, b = _rgb_of_cmyk(*vals)
    elif space == "/DeviceGray" and len(vals) == 1:
        r = g = b = vals[0]
    else:
        return S.Fill.plain(RED)
    return S.Fill.plain(S.dissolve(S.colour_of_rgba_float(r, g, b, 1.0), delta))


_JOIN = {0: abi.JOIN_MITRED, 1: abi.JOIN_ROUND, 2: abi.JOIN_BEVEL}   # render.ml:1517-1522
_CAP = {0: abi.CAP_BUTT, 1: abi.CAP_ROUND, 2: abi.CAP_PROJECTING}    # render.ml:1511-1515


def scene_of_graphic(elts):
    """Render.scene_of_graphic (render.ml:1524-1565): [(kind, fill, winding, subpaths, strokespec | None)] in the order of
    the reference's list — per path the stroked object, then the filled one."""
    out = []
    for e in elts:
        if e[0] == "Path":
            _, winding, subpaths, a = e
            segs = [list(sp[1]) for sp in subpaths]
            if a["line"] is not None:
                if a["capstyle"] not in _CAP:
                    raise PdfError("Bad PDF cap")
                if a["joinstyle"] not in _JOIN:
                    raise PdfError("bad PDF join")
                spec = (_CAP[a["capstyle"]], _JOIN[a["joinstyle"]], _CAP[a["capstyle"]], a["mitrelimit"], a["linewidth"])
                out.append(("stroke", fill_of_pdf_colour(a["line"][0], a["line"][1], a["line_transparency"]), winding, segs, spec))
            if a["fill"] is not None:
                out.append(("fill", fill_of_pdf_colour(a["fill"][0], a["fill"][1], a["fill_transparency"]), winding, segs, None))
        elif e[0] == "MCSection":
            out.extend(scene_of_graphic(e[1]))
        elif e[0] == "Clip":
            out.extend(scene_of_graphic(e[2]))
    return out


def stroke_edges_host(spec, subpaths):
    """Shapes.strokepath on the host: the product's stroker (coh_host_strokepath), its outline flattened by
    coh_host_edgelist_of_subpath, sorted by Polygon.sort_edgelist_maxy_rev (stable).  (coh_strokepath does the flattening
    and the sort on the device.)"""
    out, counts, _ = abi.host_strokepath(spec, subpaths)
    parts, k = [], 0
    for c in counts:
        segs = []
        for r in out[k : k + int(c)]:
            pts = [(float(r[1 + 2 * j]), float(r[2 + 2 * j])) for j in range(4)]
            segs.append(("L", pts[0], pts[1]) if r[0] == 0.0 else ("C", pts[0], pts[1], pts[2], pts[3]))
        parts.append(abi.host_edgelist_of_subpath(segs))
        k += int(c)
    e = np.concatenate(parts) if parts else np.zeros((0, 4), np.int32)
    if len(e):
        e = e[np.argsort(-np.maximum(e[:, 1], e[:, 3]), kind="stable")]
    return e


def add_pdf_page(b, pdf, scale=1.0, origin=(0.0, 0.0), flip_height=None, ctx=None, **group_kw):
    """Append examples.ml's load_text (158-165): Obj (Group (rev objs), transform, Over) with the transform applied to the
    geometry as render.ml:196-208 does (points mapped, line widths multiplied by |scale|): x' = ox + scale x and
    y' = oy + scale y, or oy + scale (flip_height - y) when flip_height is given (PDF user space has y upwards).
    ctx: flatten and sort the stroke outlines on the device (coh_strokepath) instead of on the host.
    Returns the number of objects in the group."""
    elts, _ = graphic_of_page(pdf)
    objs = scene_of_graphic(elts)
    if not objs:
        raise PdfError("renderobjects_of_graphic produced no content")
    ox, oy = origin

    def tr(p):
        return (ox + scale * p[0], oy + scale * ((flip_height - p[1]) if flip_height is not None else p[1]))

    def tr_segs(subpaths):
        return [[(s[0],) + tuple(tr(p) for p in s[1:]) for s in sp] for sp in subpaths]

    b.group_begin(**group_kw)
    for kind, fill, winding, subpaths, spec in reversed(objs):
        w = S.COH_NONZERO if winding == "NonZero" else S.COH_EVENODD
        sub = tr_segs(subpaths)
        if kind == "fill":
            b.path([sp for sp in sub if sp], fill, w)
        else:
            sp_ = abi.strokespec(spec[0], spec[1], spec[2], spec[3], spec[4] * abs(scale))
            edges = ctx.strokepath(sp_, sub)[0] if ctx is not None else stroke_edges_host(sp_, sub)
            b.stroked_path_edges(edges, fill)
    b.group_end()
    return len(objs)
