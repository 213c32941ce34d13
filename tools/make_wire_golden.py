"""Golden vectors of the front end's socket format, made by the REFERENCE's own Python side of it.

Imports /root/reference/pycaml.py (the marshal the wx front end speaks, pycaml.py:30-98; the OCaml side is
camlpy.ml:18-124) in this container and writes tests/golden/wire_pycaml.json: for every value the bytes
pycaml.marshall gives, and what pycaml.unmarshall reads back from them.  pycaml.py is Python 2 text; the only thing
Python 3 lacks for these two functions is the name types.BooleanType, bound to bool below — the file is not edited.

    python tools/make_wire_golden.py
"""
import importlib.util
import json
import os
import random
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
types.BooleanType = bool
spec = importlib.util.spec_from_file_location("pycaml", "/root/reference/pycaml.py")
pycaml = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pycaml)


def random_value(rng, depth=0):
    k = rng.randrange(6 if depth < 4 else 4)
    if k == 0:
        return None
    if k == 1:
        return rng.random() < 0.5
    if k == 2:
        return rng.choice([0, 1, 255, 256, 65535, 1 << 24, (1 << 31) - 1, rng.randrange(1 << 31)])
    if k == 3:
        return "".join(chr(rng.randrange(256)) for _ in range(rng.randrange(0, 40)))
    return [random_value(rng, depth + 1) for _ in range(rng.randrange(0, 5))]


def main():
    rng = random.Random(20261019)
    rgb = "".join(chr(rng.randrange(256)) for _ in range(7 * 5 * 3))
    values = [
        None, True, False, 0, 1, 255, 65536, (1 << 31) - 1, -1, -2147483648, "", "RefreshWindow",
        "".join(chr(i) for i in range(256)), [], [[]], [None], [[], [[]], []],
        ["Internal", "RefreshWindow"], ["MouseNow", 3], ["AppClose"],
        ["MakeWindow", "lion", 640, 480, 20, 30, 1280, 1024, True],                    # wxgui.ml:276-277
        ["RefreshWindow", 2, 11, 13, 7, 5, rgb],                                       # wxgui.ml:360-363
        ["Internal", "MouseNow", 412, 77], ["KeyDown", 2, 65, False, True, None],
    ] + [random_value(rng) for _ in range(60)]
    cases = []
    for v in values:
        msg = pycaml.marshall(v)                      # str of chars 0..255
        back = pycaml.unmarshall(msg[4:])             # what mltalk.py:30-45 does with the 4 size bytes taken off
        cases.append({"value": v, "hex": msg.encode("latin-1").hex(), "unmarshalled": back})
    with open(os.path.join(ROOT, "tests", "golden", "wire_pycaml.json"), "w") as f:
        json.dump({"made_by": "tools/make_wire_golden.py from /root/reference/pycaml.py (marshall, unmarshall)", "cases": cases}, f, indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
