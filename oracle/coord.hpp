// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's integer sub-pixel grid.
// Follows /root/reference/coord.ml:23-54.  PARITY UNPINNED: the reference ships no
// tests/golden vectors and cannot be built here (no OCaml toolchain).
#pragma once
#include <cmath>
namespace oracle {
// coord.ml:23-27
constexpr int ipspacing = 32;
constexpr int halfips = ipspacing / 2;
// coord.ml:34-41
inline int right_of_pix(int p) { return p * ipspacing; }
inline int left_of_pix(int p) { return right_of_pix(p) - ipspacing + 1; }
// coord.ml:44 — OCaml `/` truncates toward zero, as does C++ `/`.
inline int pix_of_sub(int n) { return (n + ipspacing - 1) / ipspacing; }
// coord.ml:47 — toint = int_of_float (truncation), after ceil.
inline int sub_of_float(double f) { return (int)std::ceil(f * 32.0 - 16.0); }
// coord.ml:50
inline int pix_of_float(double f) { return pix_of_sub(sub_of_float(f)); }
}  // namespace oracle
