"""The C ABI from plain C: tests/cpp/abi_every_symbol.c calls every exported entry point of include/coherence_b200.h
(and its error paths) with plain host buffers; and the OCaml stubs (ocaml/coherence_stubs.c) are checked against the
same header with stand-in runtime headers (there is no OCaml toolchain in this image: syntax and call signatures only)."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "cpp", "abi_every_symbol")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cpp")])


def test_c_program_builds_and_reports_missing_device():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300).stdout
    import torch

    if torch.cuda.is_available():
        assert "ABI-EVERY-SYMBOL PASS" in out, out
    else:
        assert "no CPU fallback" in out and "ABI-EVERY-SYMBOL NO-DEVICE" in out


def test_c_program_names_every_symbol_of_the_header():
    """Every function declared in the header is called by the C program (and bound by the OCaml stubs)."""
    hdr = open(os.path.join(ROOT, "include", "coherence_b200.h")).read()
    declared = set(re.findall(r"\b(coh_[a-z0-9_]+)\s*\(", hdr))
    prog = open(os.path.join(HERE, "cpp", "abi_every_symbol.c")).read()
    stubs = open(os.path.join(ROOT, "ocaml", "coherence_stubs.c")).read()
    assert not sorted(d for d in declared if d + "(" not in prog), "entry points the C test does not call"
    assert not sorted(d for d in declared if not re.search(r"\b" + d + r"\b", stubs)), "entry points the OCaml stubs do not bind"
    ml = open(os.path.join(ROOT, "ocaml", "coherence_gpu.ml")).read()
    prims = set(re.findall(r"CAMLprim value (coh_ml_[a-z0-9_]+)\(", stubs))
    assert not sorted(p for p in prims if '"' + p + '"' not in ml), "stubs without an `external` in coherence_gpu.ml"


def test_ocaml_stubs_compile_against_the_abi():
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror", "-fsyntax-only",
                           "-I" + os.path.join(HERE, "cpp", "fake_caml"), "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "ocaml", "coherence_stubs.c")])


@pytest.mark.gpu
def test_every_symbol_from_plain_c():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ABI-EVERY-SYMBOL PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
