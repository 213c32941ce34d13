"""Fixed scenes behind tests/golden/oracle_frames.json (tools/make_golden.py)."""
import math

from coherence_renderer_b200 import scene as S


def _lion():
    return S.lion_scene(640, 480, 1.4)


def _random():
    return S.random_scene(400, 300, 80, seed=0xC0FFEE, brush_fraction=0.25)


def _groups():
    b = S.SceneBuilder()
    b.polygon([(20.3, 20.1), (120.7, 30.2), (60.2, 150.9)], S.Fill.plain(S.dissolve(S.rgba8(200, 30, 30), 180)))
    b.group_begin(pretrans=140)
    b.polygon([(40.0, 40.0), (200.5, 60.5), (90.0, 180.0)], S.Fill.gradient((40.0, 40.0), (200.0, 180.0), True, True, S.rgba8(10, 200, 40), S.rgba8(0, 0, 120)))
    b.group_begin()
    b.rectangle(S.rgba8(255, 255, 0), 100.0, 20.0, 180.0, 120.0, pretrans=90)
    b.group_end()
    b.group_end()
    b.cpg("xor", [S.polygon_segments([(10.2, 40.3), (190.6, 43.1), (188.0, 100.2), (12.0, 97.7)])],
          [S.polygon_segments([(30.0, 20.4), (170.0, 70.2), (160.0, 140.9), (25.0, 72.6)])], S.Fill.plain(S.rgba8(30, 30, 200)))
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, 256.0, 200.0)
    return b


def _filters():
    b = S.SceneBuilder()
    circ = [S.polygon_segments([(100.3 + 50.5 * math.cos(2 * math.pi * i / 28), 80.2 + 50.5 * math.sin(2 * math.pi * i / 28)) for i in range(28)])]
    b.filter("blur", circ, kernel=("gaussian", 3))
    b.polygon([(30.3, 30.2), (150.5, 33.9), (148.1, 130.7), (28.8, 124.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
    b.filter("monochrome", [S.polygon_segments([(120.0, 60.0), (195.0, 66.0), (180.0, 150.0), (110.0, 140.0)])])
    b.polygon([(5.0, 5.0), (195.0, 8.0), (185.0, 150.0), (12.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)), convolve=("unit", 2))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, 200.0, 160.0)
    return b


SCENES = {
    "lion_640x480": (_lion, 640, 480, (0, 0, 640, 480)),
    "random_80_objects": (_random, 400, 300, (0, 0, 400, 300)),
    "groups_gradients_cpg": (_groups, 256, 200, (0, 0, 256, 200)),
    "groups_partial_update": (_groups, 256, 200, (40, 30, 150, 120)),
    "filters_convolved": (_filters, 200, 160, (0, 0, 200, 160)),
}
