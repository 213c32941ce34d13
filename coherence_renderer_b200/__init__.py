"""coherence_renderer_b200 — B200-native raster hot path of the Coherence 2-D renderer.

The product is the C-ABI shared library `libcoherence_b200.so` (include/coherence_b200.h):
hand-written sm_100a CUDA kernels behind the entry points the reference's OCaml modules
(render.mli, polygon.mli, fill.mli, sprite.mli) would bind.  This Python package is the
thin host-side mirror used by tests and bench.py: ctypes bindings (`abi`), scene
construction mirroring render.ml's `renderobject` constructors (`scene`), and benchmark
scenes (`scenes`).  There is no CPU fallback: everything here fails loudly when the CUDA
library or a GPU is missing.
"""
from . import abi, scene  # noqa: F401
