"""A SECOND, INDEPENDENT restatement of the reference's scan converter and compositing arithmetic, written
directly from /root/reference/polygon.ml:235-240, 314-316, 326-609, 616-705, coord.ml:23-47 and colour.ml:287-361 —
NOT from oracle/ (the C++ restatement) and not from the CUDA code.  Plain Python lists and dictionaries, the
reference's own recursion unrolled into loops, sets of pixels instead of span algebra where the reference calls
sprite.ml.  TEST INFRASTRUCTURE: tests/test_independent_restatement.py runs it differentially against the oracle
on thousands of random edge lists, so that an error of transcription in either restatement shows up as a
disagreement (the reference itself cannot be built here: no OCaml toolchain).

Conventions: an edge is a tuple (x0, y0, x1, y1) of sub-bin integers; OCaml `/` and `toint` truncate toward zero.
"""
import math

IPSPACING = 32  # coord.ml:23
HALFIPS = IPSPACING // 2  # coord.ml:27
RES = 32  # polygon.ml:22-26
SOFTNESS = 2.0


def tdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def toint(f):
    return int(f)  # truncation toward zero, like int_of_float


def pix_of_sub(n):  # coord.ml:44
    return tdiv(n + IPSPACING - 1, IPSPACING)


def left_of_pix(p):  # coord.ml:34-37
    return p * IPSPACING - IPSPACING + 1


# ---- polygon.ml:235-240: projections ------------------------------------------------------
def x0in(e):
    x0, y0, x1, y1 = e
    return x1 if y0 > y1 else (x0 if y1 > y0 else min(x0, x1))


def x1in(e):
    x0, y0, x1, y1 = e
    return x0 if y0 > y1 else (x1 if y1 > y0 else max(x0, x1))


def xminin(e):
    return min(e[0], e[2])


def xmaxin(e):
    return max(e[0], e[2])


def yminin(e):
    return min(e[1], e[3])


def ymaxin(e):
    return max(e[1], e[3])


def direction(e):  # polygon.ml:326-328 crossing_of_line: A if y1 > y0 else C; val_of_dir A = 1, C = -1 (314-316)
    return 1 if e[3] > e[1] else -1


def gradient(e):  # polygon.ml:532-535
    denom = ymaxin(e) - yminin(e)
    if denom == 0:
        return 0.0
    return float(x1in(e) - x0in(e)) / float(denom)


# ---- polygon.ml:332-388 clip_yrange2_points -------------------------------------------------
def clip_yrange_points(top, bot, active):
    """active: list of (g, edge).  Returns tops, middles, bots (crossings are (pos, dir))."""
    tops, middles, bots = [], [], []
    for g, edge in active:
        x0, x1 = x0in(edge), x1in(edge)
        ymin, ymax = yminin(edge), ymaxin(edge)
        if ymin > bot or ymax < top:
            continue
        if ymin == ymax:
            middles.append(edge)
        elif ymin >= top and ymax <= bot:
            middles.append(edge)
        elif ymin >= top:  # just bottom clipping
            y = bot
            xy = toint(float(x0) + g * (float(y - ymin) + 0.25) + 0.5)
            middles.append((x0, ymin, xy, y))
            bots.append((xy, direction(edge)))
        elif ymax <= bot:  # just top clipping
            y = top - 1
            xy = toint(float(x0) + g * (float(y - ymin) + 0.25) + 0.5)
            middles.append((xy, y + 1, x1, ymax))
            tops.append((xy, direction(edge)))
        else:  # clip both: the bottom crossing restarts from the clipped edge (polygon.ml:365-379)
            y = top - 1
            xy = toint(float(x0) + g * (float(y - ymin) + 0.25) + 0.5)
            topcrossing = xy
            edge2 = (xy, y + 1, x1, ymax)
            y = bot
            x0b, yminb = x0in(edge2), yminin(edge2)
            xyb = toint(float(x0b) + g * (float(y - yminb) + 0.25) + 0.5)
            middles.append((x0b, yminb, xyb, y))
            tops.append((topcrossing, direction(edge)))
            bots.append((xyb, direction(edge)))
    return tops, middles, bots


# ---- spans as SETS of pixels (the reference keeps sorted (start, length) lists and merges overlapping or abutting
# ones in spanacc, polygon.ml:394-403; a set of integers has the same meaning and no merging rules to get wrong) ----
def _add(pixels, s, e):
    if e >= s:
        pixels.update(range(s, e + 1))


def coverage(middles):  # polygon.ml:444-453
    px = set()
    for e in middles:
        _add(px, pix_of_sub(xminin(e) - HALFIPS), pix_of_sub(xmaxin(e) + HALFIPS))
    return px


def spans_of_edgepoints(points, aa):  # even-odd, polygon.ml:456-479: pairs (1st, 2nd), (3rd, 4th) ...
    pts = sorted(points, key=lambda c: c[0])  # stable
    px = set()
    i = 0
    while i + 1 < len(pts):
        p, q = pts[i][0], pts[i + 1][0]
        if aa:
            _add(px, pix_of_sub(p), pix_of_sub(q))
        else:
            _add(px, pix_of_sub(p - HALFIPS), pix_of_sub(q + HALFIPS))
        i += 2
    return px


def nonzero_findspans(points, aa):  # polygon.ml:482-512
    pts = sorted(points, key=lambda c: c[0])
    px = set()
    c = 0
    for i in range(len(pts) - 1):  # the last crossing has no successor: ignored
        c += pts[i][1]
        if c != 0:
            p, q = pts[i][0], pts[i + 1][0]
            if aa:
                _add(px, pix_of_sub(p), pix_of_sub(q))
            else:
                _add(px, pix_of_sub(p - HALFIPS), pix_of_sub(q + HALFIPS))
    return px


def shapeminshape_spanline(tops, middles, bots, winding, aa):  # polygon.ml:520-528
    f = nonzero_findspans if winding == 0 else spans_of_edgepoints
    t, b, c = f(tops, aa), f(bots, aa), coverage(middles)
    tbc = t | b | c
    return tbc, tbc - c


# ---- polygon.ml:538-603: the row loop --------------------------------------------------------
def shapeminshape_rows(edges, winding, aa=False):
    """Returns ({y: set of x}, {y: set of x}) for shape and minshape (rows without pixels are absent)."""
    if not edges:
        return {}, {}
    mel = sorted(edges, key=lambda e: -ymaxin(e))  # sort_edgelist_maxy_rev (stable)
    y = pix_of_sub(ymaxin(mel[0]) + HALFIPS)  # polygon.ml:564
    ael = []
    shape, minshape = {}, {}
    while True:
        top = left_of_pix(y) - HALFIPS
        bottom = top + 2 * IPSPACING - 1
        k = 0
        while k < len(mel) and ymaxin(mel[k]) >= top:  # cleavewhile_unordered: a prefix of the sorted list
            k += 1
        newly, mel_rest = mel[:k], mel[k:]
        ael = [(g, e) for g, e in ael if not (yminin(e) > bottom)] + [(gradient(e), e) for e in newly]
        if not mel and not ael:  # polygon.ml:548-550 (tests the list BEFORE this row's activation)
            break
        tops, middles, bots = clip_yrange_points(top, bottom, ael)
        s, m = shapeminshape_spanline(tops, middles, bots, winding, aa)
        if s:
            shape[y] = s
        if m:
            minshape[y] = m
        mel = mel_rest
        y -= 1
    return shape, minshape


def flat_of_rows(rows):
    """The canonical flat export used across the C ABI: per non-empty row (increasing y): y, nspans, (x, len) ..."""
    out = []
    for y in sorted(rows):
        xs = sorted(rows[y])
        spans, start, prev = [], xs[0], xs[0]
        for x in xs[1:]:
            if x != prev + 1:
                spans.append((start, prev - start + 1))
                start = x
            prev = x
        spans.append((start, prev - start + 1))
        out += [y, len(spans)]
        for s, l in spans:
            out += [s, l]
    return out


# ---- polygon.ml:616-705: antialiasing ---------------------------------------------------------
def _maintable():
    def pos(p):
        return (float(p - 1) * 6.0) / float(RES - 1) - 3.0

    t = [[0] * RES for _ in range(RES)]
    for x in range(1, RES + 1):
        for y in range(1, RES + 1):
            xp, yp = pos(x), pos(y)
            t[x - 1][y - 1] = toint(math.exp(-((xp * xp + yp * yp) / SOFTNESS)) * 255.0)
    return t


MAINTABLE = _maintable()
VOLUME = tdiv(sum(sum(r) for r in MAINTABLE) * 256, 255)  # polygon.ml:646-647


def scaled_shape_rows(edges, winding):  # polygon.ml:673-692: shape of the edge list scaled by res / 2, `_aa` span rules
    h = RES // 2
    scaled = [(x0 * h, y0 * h, x1 * h, y1 * h) for x0, y0, x1, y1 in edges]
    return shapeminshape_rows(scaled, winding, aa=True)[0]


def pixel_opacity(scaled_rows, x, y):  # polygon.ml:694-705 + 650-651
    h = RES // 2
    minx, miny = (x - 1) * h - h, (y - 1) * h - h
    count = 0
    for yy in range(miny, miny + RES):
        row = scaled_rows.get(yy)
        if not row:
            continue
        for xx in range(minx, minx + RES):
            if xx in row:
                count += MAINTABLE[xx - minx][yy - miny] * 256  # lookup_in_table over runs = sum over their pixels
    return tdiv(count + tdiv(VOLUME, 2), VOLUME)


# ---- colour.ml:287-361 on (r, g, b, a) tuples -------------------------------------------------
def div255(i):
    return (i + (i >> 8) + 1) >> 8


def dissolve(col, delta):
    assert 0 <= delta <= 255
    if delta == 0:
        return (0, 0, 0, 0)
    if delta == 255:
        return col
    return tuple(div255(c * delta) for c in col)


def prelerp(p, q, a):
    t = a * p + 128
    return p + q - (((t >> 8) + t) >> 8)


def over(a, b):
    aa = a[3]
    if aa == 0:
        return b
    if aa == 255:
        return a
    return (prelerp(b[0], a[0], aa), prelerp(b[1], a[1], aa), prelerp(b[2], a[2], aa), prelerp(b[3], aa, aa))


def alpha_over(a, b):
    aa = a[3]
    if aa == 0:
        return b
    if aa == 255:
        return a
    return (0, 0, 0, prelerp(b[3], aa, aa))


def pd_plus(a, b):
    out = tuple(x + y for x, y in zip(a, b))
    assert all(v <= 255 for v in out)
    return out


def dissolve_between(a, b, alpha):
    assert 0 <= alpha <= 255
    if alpha == 0:
        return b
    if alpha == 255:
        return a
    return pd_plus(dissolve(a, alpha), dissolve(b, 255 - alpha))
