"""N4 meeting N1 — the front end's socket format (camlpy.mli) and the RefreshWindow message (wxgui.ml:333-366).

CPU: coh_host_wire_marshal / coh_host_wire_unmarshal against golden vectors MADE BY THE REFERENCE ITSELF — its Python side
of the same format (pycaml.py, run by tools/make_wire_golden.py; tests/golden/wire_pycaml.json) — plus the behaviour of
Camlpy.unmarshall on incomplete and malformed messages (camlpy.ml:88-124), and the RefreshWindow header.
GPU: coh_wire_refresh_window: the whole message with the pixels of a rendered frame against the oracle's frame.
"""
import json
import os

import numpy as np
import pytest

from coherence_renderer_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_pycaml.json")


def _cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def _bytes_form(v):
    """The value with its strings as bytes (what unmarshal returns)."""
    if isinstance(v, str):
        return v.encode("latin-1")
    if isinstance(v, list):
        return [_bytes_form(e) for e in v]
    return v


def test_oracle_is_pinned_by_the_reference_python_side(oracle):
    """oracle/camlpy.hpp (the list-based restatement of camlpy.ml) against the vectors pycaml.py produced: this part of the
    oracle is pinned by the reference's own code before it checks the product."""
    for c in _cases():
        msg = bytes.fromhex(c["hex"])
        assert oracle.wire_marshal(c["value"]) == msg, c["value"]
        assert oracle.wire_unmarshal(msg) == (len(msg), _bytes_form(c["unmarshalled"]))
        assert oracle.wire_unmarshal(msg[:-1]) is None


def test_marshal_equals_the_reference_python_side():
    cases = _cases()
    assert len(cases) >= 80
    for c in cases:
        assert abi.host_wire_marshal(c["value"]).hex() == c["hex"], c["value"]


def test_unmarshal_equals_the_reference_python_side():
    for c in _cases():
        msg = bytes.fromhex(c["hex"])
        taken, v = abi.host_wire_unmarshal(msg)
        assert taken == len(msg)
        assert v == _bytes_form(c["unmarshalled"]), c["value"]
        # Ints come back without sign extension (camlpy.ml:85-86), everything else round-trips
        if not (isinstance(c["value"], int) and not isinstance(c["value"], bool) and c["value"] < 0):
            assert v == _bytes_form(c["value"])
        # a second message behind it is left alone (camlpy.ml:113: String.sub str 4 len)
        taken2, v2 = abi.host_wire_unmarshal(msg + b"\x00\x00\x00\x01\x01")
        assert taken2 == len(msg) and v2 == v


def test_unmarshal_waits_for_the_whole_message():
    msg = abi.host_wire_marshal(["RefreshWindow", 2, 11, 13, 7, 5, bytes(range(105))])
    for n in range(len(msg)):
        assert abi.host_wire_unmarshal(msg[:n]) is None       # camlpy.ml:108, 111: None
    assert abi.host_wire_unmarshal(msg)[0] == len(msg)


def _framed(body):
    return len(body).to_bytes(4, "big") + body


@pytest.mark.parametrize("body", [
    b"",                                      # no value at all (camlpy.ml:119-121 wants [x])
    b"\x01\x01",                              # two values
    b"\x05",                                  # unknown tag
    b"\x02\x00\x00\x01",                      # Int cut short
    b"\x04",                                  # Bool cut short
    b"\x03\x00\x00\x00\x05abcd",              # String longer than what is left
    b"\x00\x00\x00\x00\x06\x02\x00\x00\x00\x01",      # Tuple longer than what is left
    b"\x00\x00\x00\x00\x03\x02\x00\x00\x00\x01\x01",  # a member that runs over its Tuple's end
    b"\x00\x00\x00\x00\x07\x03\x00\x00\x00\x09ab",     # String inside a Tuple longer than the Tuple
    b"\x00\x00\x00",                          # Tuple header cut short
])
def test_unmarshal_invalid_data(body, oracle):
    with pytest.raises(abi.CohError):
        abi.host_wire_unmarshal(_framed(body))
    with pytest.raises(oracle.OracleError):                   # the oracle's list-based unmarshall_inner agrees
        oracle.wire_unmarshal(_framed(body))


def test_unmarshal_bool_and_nesting():
    assert abi.host_wire_unmarshal(_framed(b"\x04\x07"))[1] is True            # camlpy.ml:94: b <> 0
    deep = None
    for _ in range(3000):                                                      # no recursion in the library
        deep = [deep]
    msg = abi.host_wire_marshal_tokens([0] * 3000 + [1], [1] * 3000 + [0], [0] * 3001)
    assert msg is not None and len(msg) == 4 + 5 * 3000 + 1
    nt = abi.C.c_int32()
    tk = abi.C.c_int64()
    buf = np.frombuffer(msg, dtype=np.uint8)
    assert abi.lib().coh_host_wire_unmarshal(buf.ctypes.data_as(abi.C.POINTER(abi.C.c_uint8)), abi.C.c_int64(len(msg)), None, None, None, 0,
                                             abi.C.byref(nt), abi.C.byref(tk)) == 0
    assert nt.value == 3001 and tk.value == len(msg)


def test_marshal_rejects_token_lists_that_are_not_one_value():
    T, U, I = abi.WIRE_TUPLE, abi.WIRE_UNIT, abi.WIRE_INT
    assert abi.host_wire_marshal_tokens([U, U], [0, 0], [0, 0]) is None        # two values
    assert abi.host_wire_marshal_tokens([T, I], [2, 5], [0, 0]) is None        # a member missing
    assert abi.host_wire_marshal_tokens([9], [0], [0]) is None                 # no such tag
    assert abi.host_wire_marshal_tokens([T, T, I, U], [2, 1, 5, 0], [0] * 4).hex() == "00000010" + "000000000b" + "0000000005" + "0200000005" + "01"


def test_refresh_window_header():
    rgb = bytes((i * 7) & 255 for i in range(7 * 5 * 3))
    total, hdr = abi.host_wire_refresh_window(2, 11, 13, 17, 17)               # xmin, ymin, xmax, ymax INCLUSIVE: 7 x 5
    want = abi.host_wire_marshal(["RefreshWindow", 2, 11, 13, 7, 5, rgb])
    assert total == len(want) and hdr + rgb == want
    # the golden RefreshWindow message of the reference's Python side has the same frame
    gold = [c for c in _cases() if isinstance(c["value"], list) and c["value"][:1] == ["RefreshWindow"]][0]
    assert bytes.fromhex(gold["hex"])[: len(hdr)] == hdr
    # wxgui.ml:357: zero-width or zero-height rectangles just do nothing
    assert abi.host_wire_refresh_window(1, 5, 5, 5, 9) == (0, b"") and abi.host_wire_refresh_window(1, 5, 9, 8, 9) == (0, b"")
    for bad in [(1, 9, 5, 5, 9), (1, 5, 9, 8, 5), (1, -1, 0, 4, 4)]:          # wxgui.ml:335-336
        with pytest.raises(abi.CohError):
            abi.host_wire_refresh_window(*bad)


@pytest.mark.gpu
def test_wire_refresh_window_against_oracle(ctx, oracle):
    """The message a front end receives for a dirty rectangle: framing as Camlpy.marshall makes it, pixels = the
    premultiplied r, g, b bytes Wxgui.plot_sprite writes (wxgui.ml:368-375) of the oracle's frame."""
    from coherence_renderer_b200 import scene as S

    W, H = 320, 240
    b = S.lion_scene(W, H, 0.7)
    b.polygon([(30.5, 20.2), (290.1, 60.7), (120.9, 220.3)], S.Fill.plain(S.dissolve(S.rgba8(30, 60, 220), 120)))
    objs, n, nbg, edges, points = b.arrays()
    ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    for (x0, y0, x1, y1) in [(37, 21, 237, 170), (0, 0, W - 1, H - 1), (100, 100, 101, 101), (W - 2, 0, W - 1, H - 1)]:
        msg = ctx.wire_refresh_window(3, x0, y0, x1, y1)
        taken, v = abi.host_wire_unmarshal(msg)
        w, h = x1 - x0 + 1, y1 - y0 + 1
        assert taken == len(msg) and v[:6] == [b"RefreshWindow", 3, x0, y0, w, h]
        want = ref[y0 : y1 + 1, x0 : x1 + 1]
        rgb = np.stack([(want & 255), (want >> 8) & 255, (want >> 16) & 255], axis=-1).astype(np.uint8)
        assert v[6] == rgb.tobytes()
    assert ctx.wire_refresh_window(3, 10, 10, 10, 50) == b""                   # nothing to send
    with pytest.raises(abi.CohError):
        ctx.wire_refresh_window(3, 10, 10, W, 50)                              # outside the framebuffer
    ctx.scene_free(sc)


def _camlpy_marshall(m):
    """camlpy.ml:39-82 restated in the test (big-endian low 32 bits, tags 0..4), independent of the library."""
    def flat(v):
        if v is None:
            return b"\x01"
        if isinstance(v, bool):
            return b"\x04" + (b"\x01" if v else b"\x00")
        if isinstance(v, int):
            return b"\x02" + (v & 0xFFFFFFFF).to_bytes(4, "big")
        if isinstance(v, bytes):
            return b"\x03" + len(v).to_bytes(4, "big") + v
        inner = b"".join(flat(e) for e in v)
        return b"\x00" + len(inner).to_bytes(4, "big") + inner
    body = flat(m)
    return len(body).to_bytes(4, "big") + body


def test_round_trips_of_random_values(oracle):
    from hypothesis import given, settings, strategies as st

    pyoracle = oracle

    leaves = st.one_of(st.none(), st.booleans(), st.integers(min_value=0, max_value=2**32 - 1), st.binary(max_size=60))
    values = st.recursive(leaves, lambda inner: st.lists(inner, max_size=6), max_leaves=40)

    @settings(max_examples=300, deadline=None)
    @given(values)
    def check(v):
        msg = abi.host_wire_marshal(v)
        assert msg == _camlpy_marshall(v) == pyoracle.wire_marshal(v)
        assert pyoracle.wire_unmarshal(msg) == (len(msg), v)
        taken, back = abi.host_wire_unmarshal(msg)
        assert taken == len(msg) and back == v
        for cut in (1, len(msg) // 2, len(msg) - 1):
            assert abi.host_wire_unmarshal(msg[:cut]) is None

    check()
