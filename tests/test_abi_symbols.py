"""The C-ABI library loads, exports every symbol include/coherence_b200.h declares, and fails loudly without a
GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from coherence_renderer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "coherence_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(coh_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert _declared() == sorted(abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(abi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} is declared in coherence_b200.h but not exported"


def test_struct_layout_matches_header(tmp_path):
    """Compile the header with gcc and compare sizeof/offsetof with the ctypes mirror."""
    import subprocess

    fields = [f[0] for f in abi.CohObject._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "coherence_b200.h"\nint main(){printf("%zu", sizeof(coh_object));' + "".join(
        f'printf(" %zu", offsetof(coh_object, {f}));' for f in fields) + "return 0;}"
    src = tmp_path / "layout.c"
    src.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(abi.CohObject)
    assert out[1:] == [getattr(abi.CohObject, f).offset for f in fields]


def test_colour_codec_is_the_reference_encoding(oracle):
    lib = abi.lib()
    for w in (0, 0xFFFFFFFF, 0xFF000000, 0xFF0000FF, (200 << 24) | (30 << 16) | (20 << 8) | 10, 0x80402010, 0x7F7F7F7F, 0x01010101):
        c = lib.coh_colour_of_rgba8(C.c_uint32(w))
        assert c == oracle.colour_of_rgba8(w)
        assert lib.coh_rgba8_of_colour(C.c_int32(c)) == w


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(abi.CohError, match="no CUDA device"):
        abi.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "coherence_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle/", "ORACLE_DIR_MENTION") or "import" not in text or "pyoracle" not in text, f
                assert "pyoracle" not in text and "liboracle" not in text, f"{f} references the oracle"
