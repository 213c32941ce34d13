// raster_core.cuh — the per-row arithmetic of the B200 raster hot path.
//
// Everything here is written from the specification in SURVEY.md Appendix A (itself a
// restatement of /root/reference/{coord,colour,polygon,fill}.ml) and is shaped for the
// GPU: a row band of one object is evaluated by ONE lane into bit-rows (one bit per
// pixel) instead of the reference's sorted span lists, because a canonical span list is
// just the run-length view of a pixel set and set algebra on bit-rows is word-wise
// AND/OR/ANDNOT.  Span lists only exist at the C-ABI boundary and in the HBM-resident
// cache (see spans.cuh).
//
// The functions are __host__ __device__ so that tests/ can compile this header with g++
// and drive the very same row arithmetic on the CPU against the oracle (no GPU in the
// build container).  The product library only ever calls them from kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define COH_HD __host__ __device__ __forceinline__
// big, cold or many-call-site routines stay out of line: the fused walker must fit the
// instruction cache (a 135 KB kernel stalled 37 % of its issue slots on instruction fetch)
#define COH_HD_NOINLINE __host__ __device__ __noinline__
#else
#define COH_HD inline
#define COH_HD_NOINLINE inline
#endif

namespace coh {

// ---------------------------------------------------------------------------------
// FP64 with the reference's rounding: OCaml floats are IEEE binary64 with one rounding
// per operation and no fused multiply-add (SURVEY.md §7 "hard parts").
// ---------------------------------------------------------------------------------
COH_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
COH_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
COH_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b; return r;
#endif
}

// ---------------------------------------------------------------------------------
// Coord (coord.ml:23-47).  Integer division truncates toward zero, as in OCaml.
// ---------------------------------------------------------------------------------
COH_HD int pix_of_sub(int n) { return (n + 31) / 32; }
COH_HD int imin(int a, int b) { return a < b ? a : b; }
COH_HD int imax(int a, int b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------
// Colour on decoded RGBA8 words r | g<<8 | b<<16 | a<<24 (colour.ml:287-365).  The
// 31-bit codec of the reference is lossless on premultiplied values, so operating on
// the decoded channels gives identical results; the codec lives at the ABI boundary.
// ---------------------------------------------------------------------------------
COH_HD uint32_t div255(uint32_t i) { return (i + (i >> 8) + 1) >> 8; }  // colour.ml:287
COH_HD uint32_t px_alpha(uint32_t c) { return c >> 24; }
// Two channels per 32-bit multiply: the products of 8-bit values stay inside their 16-bit halves, and so does
// every intermediate of div255 / prelerp, so the packed forms are exact.  div255 (c * 0) = 0 and div255 (c * 255) = c,
// so the reference's early-outs for delta = 0 / 255 need no branch.
COH_HD uint32_t px_dissolve(uint32_t c, int delta) {  // colour.ml:291-304
  const uint32_t d = (uint32_t)delta;
  uint32_t rb = (c & 0x00FF00FFu) * d, ga = ((c >> 8) & 0x00FF00FFu) * d;
  rb = ((rb + ((rb >> 8) & 0x00FF00FFu) + 0x00010001u) >> 8) & 0x00FF00FFu;
  ga = (ga + ((ga >> 8) & 0x00FF00FFu) + 0x00010001u) & 0xFF00FF00u;
  return rb | ga;
}
COH_HD uint32_t prelerp(uint32_t p, uint32_t q, uint32_t a) {  // colour.ml:310-311
  uint32_t t = a * p + 128u;
  return p + q - (((t >> 8) + t) >> 8);
}
// colour.ml:314-328: `over a b`, a nearer the viewer (premultiplied colours: every channel of the result fits its byte).
COH_HD uint32_t px_over(uint32_t a, uint32_t b) {
  const uint32_t aa = a >> 24;
  const uint32_t prb = b & 0x00FF00FFu, pga = (b >> 8) & 0x00FF00FFu;
  const uint32_t trb = prb * aa + 0x00800080u, tga = pga * aa + 0x00800080u;
  const uint32_t xrb = ((((trb >> 8) & 0x00FF00FFu) + trb) >> 8) & 0x00FF00FFu;
  const uint32_t xga = ((((tga >> 8) & 0x00FF00FFu) + tga) >> 8) & 0x00FF00FFu;
  const uint32_t res = (prb + (a & 0x00FF00FFu) - xrb) | ((pga + ((a >> 8) & 0x00FF00FFu) - xga) << 8);
  return aa == 0u ? b : (aa == 255u ? a : res);
}
// colour.ml:332-336 on the alpha channel alone (brush stamping only reads alpha back).
COH_HD uint32_t alpha_over(uint32_t aa, uint32_t ab) {
  if (aa == 0) return ab;
  if (aa == 255) return aa;
  return prelerp(ab, aa, aa);
}
COH_HD uint32_t px_plus(uint32_t a, uint32_t b) { return a + b; }  // colour.ml:339-352 (no channel overflows by contract)
COH_HD uint32_t px_dissolve_between(uint32_t a, uint32_t b, int alpha) {  // colour.ml:355-361
  if (alpha == 0) return b;
  if (alpha == 255) return a;
  return px_plus(px_dissolve(a, alpha), px_dissolve(b, 255 - alpha));
}

// ---------------------------------------------------------------------------------
// Prepared edge (one per input edge, built once by k_prep_edges).
// polygon.ml:235-240 (x0in/x1in/ymin/ymax), 532-535 (gradient), 326-328 (direction).
// ---------------------------------------------------------------------------------
struct EdgeRec {
  int x0in, x1in, ymin, ymax;  // sub-bins; x0in is the x at the ymin end
  double g;                    // float (x1in - x0in) /. float (ymax - ymin), 0 when horizontal
  int dir;                     // +1 if y1 > y0 (A) else -1 (C)
  int pad;
};
COH_HD EdgeRec make_edge(int x0, int y0, int x1, int y1) {
  EdgeRec e;
  if (y0 > y1) { e.x0in = x1; e.x1in = x0; }
  else if (y1 > y0) { e.x0in = x0; e.x1in = x1; }
  else { e.x0in = imin(x0, x1); e.x1in = imax(x0, x1); }
  e.ymin = imin(y0, y1); e.ymax = imax(y0, y1);
  int denom = e.ymax - e.ymin;
  e.g = denom == 0 ? 0.0 : ddiv((double)(e.x1in - e.x0in), (double)denom);
  e.dir = y1 > y0 ? 1 : -1;
  e.pad = 0;
  return e;
}
// polygon.ml:347-348: toint (float x0 +. g *. (float (y - ymin) +. 0.25) +. 0.5)
COH_HD int crossing_x(int x0, double g, int dy) {
  return (int)dadd(dadd((double)x0, dmul(g, dadd((double)dy, 0.25))), 0.5);
}

// ---------------------------------------------------------------------------------
// Bit-rows.  A bit-row covers pixels [wx0, wx0 + 32*NW); bit i of word w is pixel
// wx0 + 32*w + i.
// ---------------------------------------------------------------------------------
template <int NW>
struct BitRow {
  uint32_t w[NW];
  COH_HD void clear() {
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0u;
  }
};
// OR pixel interval [a, b] (inclusive, absolute pixel coords) into a bit-row stored at
// `bits` with stride `stride` words between consecutive words (stride lets 32 lanes
// keep private rows in shared memory without bank conflicts).
COH_HD_NOINLINE void or_interval(uint32_t* bits, int stride, int nwords, int wx0, int a, int b) {
  a -= wx0; b -= wx0;
  int nb = nwords * 32;
  if (b < 0 || a >= nb) return;
  if (a < 0) a = 0;
  if (b > nb - 1) b = nb - 1;
  if (a > b) return;
  int wa = a >> 5, wb = b >> 5;
  uint32_t ma = 0xFFFFFFFFu << (a & 31), mb = 0xFFFFFFFFu >> (31 - (b & 31));
  if (wa == wb) { bits[wa * stride] |= (ma & mb); return; }
  bits[wa * stride] |= ma;
  for (int w = wa + 1; w < wb; w++) bits[w * stride] = 0xFFFFFFFFu;
  bits[wb * stride] |= mb;
}
// single-word variant returning the mask
COH_HD uint32_t interval_mask32(int wx0, int a, int b) {
  a -= wx0; b -= wx0;
  if (b < 0 || a > 31) return 0u;
  if (a < 0) a = 0;
  if (b > 31) b = 31;
  if (a > b) return 0u;
  return (0xFFFFFFFFu << a) & (0xFFFFFFFFu >> (31 - b));
}

// ---------------------------------------------------------------------------------
// Row scan (polygon.ml:332-528 for one row band), restricted to a pixel window.
//
// For pixel row `y` of the edge list scaled by `s` (1: shape/minshape; 16: the x16
// super-sampled shape of polygon.ml:673-692) compute, inside the window [wlo, whi],
//     T ∪ B  = winding spans of the top / bottom band crossings,
//     C      = coverage spans of the clipped "middle" pieces,
// and hand every interval to the sink: sink.span(a, b) for T/B spans, sink.cover(a, b)
// for C spans (pixel coordinates, inclusive; the sink clips to its own storage).
// shape_row = T∪B∪C, minshape_row = shape_row − C (polygon.ml:520-528); the union/fuse
// of the reference's span lists is the OR of the bit-rows, and ties between equal
// crossings cannot change the pixel set, so crossings are ranked by
// (pos, list index).
//
// Windowed winding.  Sorted by position the crossings of a band split into those whose
// pixel reach ends left of the window, those that touch it, and those beyond it.  Only
// the middle group needs ordering.  The left group contributes its winding sum (and its
// parity) and acts as ONE predecessor for the first crossing that is not left; the right
// group only matters as "there is a successor", which clips to the window end.  So a
// lane keeps at most the crossings that touch its window (one or two for a 32-pixel
// tile) however complex the object is.
//
// MAXX bounds the crossings kept per list; on overflow the function returns false and
// the caller reports the object as too complex for the window (loudly).
// ---------------------------------------------------------------------------------
#ifndef COH_MAXX
#define COH_MAXX 64
#endif

// CAP = COH_MAXX: general list (dynamically indexed, lives in local memory).
// Small CAP with REG = true: every access is a fully unrolled compare-select, so the list stays
// in registers; callers check `n > CAP` and fall back to the general list (warp-uniformly).
template <int CAP, bool REG>
struct CrossListT {
  int v[CAP];  // (pos << 1) | (dir > 0), crossings touching the window
  int n;
  int cnt_left;     // sum of directions of the crossings left of the window
  int n_left;       // how many there are
  bool has_right;   // a crossing lies beyond the window
  COH_HD void init() { n = 0; cnt_left = 0; n_left = 0; has_right = false; }
  // reach of a crossing at p: [pix(p-16), pix(p+16)] (pixel spans are widened by half a
  // pixel, polygon.ml:458-459,490-491) or [pix(p), pix(p)] for the AA variant (469-479,498-512)
  COH_HD void add(bool pred, int p, int dir, bool aa, int wlo, int whi) {
    // branch-free: the 32 lanes of an AA scan sit on different rows and would otherwise diverge
    const int rhi = aa ? pix_of_sub(p) : pix_of_sub(p + 16);
    const int rlo = aa ? rhi : pix_of_sub(p - 16);
    const bool left = pred && rhi < wlo, right = pred && rlo > whi, inside = pred && !(rhi < wlo) && !(rlo > whi);
    cnt_left += left ? dir : 0;
    n_left += left ? 1 : 0;
    has_right = has_right || right;
    const int val = (p << 1) | (dir > 0);
    if (REG) {
#pragma unroll
      for (int k = 0; k < CAP; k++) v[k] = (inside && n == k) ? val : v[k];
    } else if (inside && n < CAP) v[n] = val;
    n += inside ? 1 : 0;
  }
  COH_HD int get(int i) const {
    if (!REG) return v[i];
    int r = v[0];
#pragma unroll
    for (int k = 1; k < CAP; k++) r = (i == k) ? v[k] : r;
    return r;
  }
};
typedef CrossListT<COH_MAXX, false> CrossList;

template <int CAP, bool REG, class Sink>
COH_HD void winding_spans_impl(const CrossListT<CAP, REG>& L, int winding, bool aa, int wlo, int whi, Sink& sink) {
  const int n = L.n < CAP ? L.n : CAP;
  // the left group as one predecessor: its span runs from left of the window to the first
  // crossing that is not left
  if (L.n_left > 0 && (n > 0 || L.has_right)) {
    bool emit = winding == 0 ? (L.cnt_left != 0) : ((L.n_left & 1) == 1);
    if (emit) {
      int b = whi;
      if (n > 0) {
        int first = L.get(0) >> 1;
        for (int j = 1; j < n; j++) { int pj = L.get(j) >> 1; if (pj < first) first = pj; }
        b = aa ? pix_of_sub(first) : pix_of_sub(first + 16);
      }
      sink.span(wlo, b);
    }
  }
  // (register lists are read through get(): compare-select chains, loops stay rolled to keep the
  // walker's code small enough for the instruction cache)
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    const int vi = L.get(i);
    const int pi = vi >> 1;
    // rank / running winding count of crossing i in the (pos, index) order, and its successor
    int cnt = L.cnt_left, rank = L.n_left;
    int succ = 0x7FFFFFFF; bool has = false;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      const int vj = L.get(j);
      const int pj = vj >> 1;
      const bool before_or_self = (pj < pi) || (pj == pi && j <= i);
      if (before_or_self) { cnt += (vj & 1) ? 1 : -1; rank += (j != i); }
      else { if (pj < succ) succ = pj; has = true; }
    }
    if (!has && !L.has_right) continue;  // the last crossing has no successor (polygon.ml:484, 458)
    const bool emit = winding == 0 ? (cnt != 0) : ((rank & 1) == 0);
    if (!emit) continue;
    const int a = aa ? pix_of_sub(pi) : pix_of_sub(pi - 16);
    const int b = has ? (aa ? pix_of_sub(succ) : pix_of_sub(succ + 16)) : whi;
    sink.span(a, b);
  }
}
template <class Sink>
COH_HD_NOINLINE void winding_spans(const CrossList& L, int winding, bool aa, int wlo, int whi, Sink& sink) {
  winding_spans_impl(L, winding, aa, wlo, whi, sink);
}
template <int CAP, class Sink>
COH_HD void winding_spans(const CrossListT<CAP, true>& L, int winding, bool aa, int wlo, int whi, Sink& sink) {
  winding_spans_impl(L, winding, aa, wlo, whi, sink);
}

// The scan of one row band is split into begin / per-edge / finish so that callers can feed
// edges from wherever they keep them (global memory lists, or a shared-memory stage filled
// cooperatively by a warp).
template <int CAP, bool REG>
struct ScanStateT {
  CrossListT<CAP, REG> tops, bots;
  int top, bot;     // band of the row in (scaled) sub-bins, polygon.ml:539-540
  int wlo, whi;     // window in (scaled) pixels
  bool aa;
  COH_HD bool overflow() const { return tops.n > CAP || bots.n > CAP; }
};
typedef ScanStateT<COH_MAXX, false> ScanState;
template <class State>
COH_HD void scan_begin(State& st, int y, bool aa, int wlo, int whi) {
  st.top = 32 * y - 47;  // left_of_pix y - halfips
  st.bot = st.top + 63;
  st.wlo = wlo; st.whi = whi; st.aa = aa;
  st.tops.init(); st.bots.init();
}
// Where an edge with (scaled) x extremes x0, x1 lies relative to a window: 1 = entirely left,
// 2 = entirely right, 0 = may touch it.  "Entirely" includes the few sub-bins a rounded
// crossing can leave the edge's x range by and the half-pixel widening of spans and coverage.
COH_HD int edge_side(int x0, int x1, int wlo, int whi) {
  const int exlo = imin(x0, x1) - 4, exhi = imax(x0, x1) + 4;
  if (pix_of_sub(exhi + 16) < wlo) return 1;
  if (pix_of_sub(exlo - 16) > whi) return 2;
  return 0;
}
// One edge, coordinates already scaled (polygon.ml:332-388, 444-453).  Written without
// data-dependent branches apart from the (warp-uniform) `side` test: the four clipping cases of
// the reference differ only in which crossings exist and which end points bound the middle piece.
template <class State, class Sink>
COH_HD void scan_edge(State& st, int x0, int x1, int ymin, int ymax, double g, int dir, int side, Sink& sink) {
  const int top = st.top, bot = st.bot;
  const bool active = !(ymin > bot || ymax < top);                                   // polygon.ml:338
  const bool middle_only = ymin == ymax || (ymin >= top && ymax <= bot);             // polygon.ml:340-343
  const bool cross_top = active && !middle_only && ymin < top;   // contributes a top crossing
  const bool cross_bot = active && !middle_only && ymax > bot;   // contributes a bottom crossing
  // An edge entirely to one side of the window needs no arithmetic: its crossings only count as
  // "left of the window" or "beyond it", and its coverage is outside.
  if (side == 1) {
    st.tops.cnt_left += cross_top ? dir : 0; st.tops.n_left += cross_top ? 1 : 0;
    st.bots.cnt_left += cross_bot ? dir : 0; st.bots.n_left += cross_bot ? 1 : 0;
    return;
  }
  if (side == 2) {
    st.tops.has_right = st.tops.has_right || cross_top;
    st.bots.has_right = st.bots.has_right || cross_bot;
    return;
  }
  // top crossing (polygon.ml:355-364), then the bottom crossing, which restarts from the rounded
  // top crossing when the edge is clipped at both ends (polygon.ml:365-385) and from x0 otherwise
  const int xt = crossing_x(x0, g, top - 1 - ymin);
  const int xs = cross_top ? xt : x0;
  const int xb = crossing_x(xs, g, cross_top ? (bot - top) : (bot - ymin));
  st.tops.add(cross_top, xt, dir, st.aa, st.wlo, st.whi);
  st.bots.add(cross_bot, xb, dir, st.aa, st.wlo, st.whi);
  // the clipped middle piece runs from (cross_top ? xt : x0) to (cross_bot ? xb : x1)
  const int pe = cross_bot ? xb : x1;
  const int lo = imin(xs, pe), hi = imax(xs, pe);
  const int ca = pix_of_sub(lo - 16), cb = pix_of_sub(hi + 16);  // polygon.ml:444-453
  if (active && cb >= st.wlo && ca <= st.whi) sink.cover(ca, cb);
}
template <class State, class Sink>
COH_HD bool scan_finish(const State& st, int winding, Sink& sink) {
  if (st.overflow()) return false;
  winding_spans(st.tops, winding, st.aa, st.wlo, st.whi, sink);
  winding_spans(st.bots, winding, st.aa, st.wlo, st.whi, sink);
  return true;
}

// `idx` (optional) lists the candidate edges of this row (K1 edge binning): any superset
// of the edges whose y range meets the band gives the same result, since every edge is
// re-tested against the band here.
template <class Sink>
COH_HD_NOINLINE bool scan_row(const EdgeRec* __restrict__ edges, const int* __restrict__ idx, int n_cand, int s, int y,
                              int winding, bool aa, int wlo, int whi, Sink& sink) {
  ScanState st;
  scan_begin(st, y, aa, wlo, whi);
  for (int i = 0; i < n_cand; i++) {
    const EdgeRec e = edges[idx ? idx[i] : i];
    const int x0 = e.x0in * s, x1 = e.x1in * s;
    scan_edge(st, x0, x1, e.ymin * s, e.ymax * s, e.g, e.dir, edge_side(x0, x1, wlo, whi), sink);
  }
  return scan_finish(st, winding, sink);
}

// Sinks -----------------------------------------------------------------------------
// One 32-pixel word, registers only (pixel rows of the tile walker).
struct Sink32 {
  int wx0;
  uint32_t S, C;
  COH_HD void span(int a, int b) { S |= interval_mask32(wx0, a, b); }
  COH_HD void cover(int a, int b) { uint32_t m = interval_mask32(wx0, a, b); S |= m; C |= m; }
};
// Multi-word row in SHARED memory (the AA rows of the walker), addressed by its 32-bit shared-window
// address so that one out-of-line routine serves every call site with ld/st.shared (a generic
// pointer would cost generic loads/stores, inlining it five times cost 465 instructions of I-cache).
#if defined(__CUDACC__)
#ifdef COH_ROWPUT_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void smem_row_put(uint32_t saddr, int nwords, int wx0, int a, int b) {
  a -= wx0; b -= wx0;
  const int nb = nwords * 32;
  if (b < 0 || a >= nb) return;
  if (a < 0) a = 0;
  if (b > nb - 1) b = nb - 1;
  if (a > b) return;
  const int wa = a >> 5, wb = b >> 5;
  const uint32_t ma = 0xFFFFFFFFu << (a & 31), mb = 0xFFFFFFFFu >> (31 - (b & 31));
  uint32_t v;
  if (wa == wb) {
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr + 4 * wa));
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr + 4 * wa), "r"(v | (ma & mb)));
    return;
  }
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr + 4 * wa));
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr + 4 * wa), "r"(v | ma));
  for (int w = wa + 1; w < wb; w++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr + 4 * w), "r"(0xFFFFFFFFu));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr + 4 * wb));
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr + 4 * wb), "r"(v | mb));
}
struct SinkRow {
  int wx0, nwords;
  uint32_t saddr;  // shared-window address of the row
  __device__ __forceinline__ void span(int a, int b) { smem_row_put(saddr, nwords, wx0, a, b); }
  __device__ __forceinline__ void cover(int a, int b) { smem_row_put(saddr, nwords, wx0, a, b); }
};
#else
struct SinkRow {  // host build of the same sink (tests/host_emul)
  int wx0, nwords;
  uint32_t* S;
  void put(int a, int b) {
    a -= wx0; b -= wx0;
    const int nb = nwords * 32;
    if (b < 0 || a >= nb) return;
    if (a < 0) a = 0;
    if (b > nb - 1) b = nb - 1;
    if (a > b) return;
    const int wa = a >> 5, wb = b >> 5;
    const uint32_t ma = 0xFFFFFFFFu << (a & 31), mb = 0xFFFFFFFFu >> (31 - (b & 31));
    if (wa == wb) { S[wa] |= (ma & mb); return; }
    S[wa] |= ma;
    for (int w = wa + 1; w < wb; w++) S[w] = 0xFFFFFFFFu;
    S[wb] |= mb;
  }
  void span(int a, int b) { put(a, b); }
  void cover(int a, int b) { put(a, b); }
};
#endif
// Multi-word rows in memory (export kernels, AA rows).  C may be null (AA needs only S).
struct SinkMem {
  int wx0, nwords, stride;
  uint32_t* S;
  uint32_t* C;
  COH_HD void span(int a, int b) { or_interval(S, stride, nwords, wx0, a, b); }
  COH_HD void cover(int a, int b) {
    or_interval(S, stride, nwords, wx0, a, b);
    if (C) or_interval(C, stride, nwords, wx0, a, b);
  }
};

// ---------------------------------------------------------------------------------
// Antialiasing (polygon.ml:616-705).  AA_PREFIX[j][k] = sum_{i<k} maintable[i][j]; the
// coverage of scaled row j inside a pixel's 32-wide window with occupancy mask m is the
// sum over the runs of m of AA_PREFIX[j][end+1] - AA_PREFIX[j][start].
// ---------------------------------------------------------------------------------
struct AATable {
  int prefix[32][33];
  int volume;  // polygon.ml:646-647
};
COH_HD int aa_row_sum(const int* prefix_row /*33 ints*/, uint32_t m) {
  int sum = 0;
  while (m) {
#if defined(__CUDA_ARCH__)
    int s = __ffs((int)m) - 1;
#else
    int s = __builtin_ctz(m);
#endif
    uint32_t t = ~(m >> s);  // zeros where the run continues (bits above 31-s read as 1)
#if defined(__CUDA_ARCH__)
    int l = t ? (__ffs((int)t) - 1) : (32 - s);
#else
    int l = t ? __builtin_ctz(t) : (32 - s);
#endif
    if (l > 32 - s) l = 32 - s;
    sum += prefix_row[s + l] - prefix_row[s];
    uint32_t run = (l >= 32) ? 0xFFFFFFFFu : (((1u << l) - 1u) << s);
    m &= ~run;
  }
  return sum;
}
// polygon.ml:646-647: volume = (sum of maintable * 256) / 255 = 42285 for the reference's table; the
// library checks the table it builds against this constant at start-up (division by a constant is a
// multiply-high instead of ~30 instructions)
constexpr int AA_VOLUME = 42285;
COH_HD int aa_opacity(int table_sum, int volume) {  // polygon.ml:650-651 with cov = 256 * sum
#ifdef COH_RUNTIME_VOLUME
  return (256 * table_sum + volume / 2) / volume;
#else
  (void)volume;
  return (256 * table_sum + AA_VOLUME / 2) / AA_VOLUME;
#endif
}

// ---------------------------------------------------------------------------------
// Antialiasing, interval form.  Inside the columns a pixel's window can read, one super-sampled row of an
// object is nearly always ONE run (everything left of an edge, everything right of it, the strip between two
// edges of a thin feature).  AaScan evaluates a row of the x16 edge list (polygon.ml:673-692, the `_aa` span
// rules of 469-512) without any bit-row: crossings that touch the window are kept as the two smallest of a
// sorted pair per band line, the coverage pieces as their hull plus at most one gap, and finish() either proves
// that T ∪ B ∪ C is the interval [lo, hi] minus at most one gap [glo, ghi] (two disjoint parts of the object in
// the window) or reports `complex` (more than two crossings on a band line inside the window, a second gap) — the
// caller then takes the general bit-row path.  The proof: every winding span starts / ends at the window border
// or at a crossing, and a crossing lies inside its own edge's coverage piece, so every span hangs on a coverage
// piece and either stays on its side of the gap or reaches across all of it; with one coverage component
// everything is connected, with two the union is the hull, or the hull minus the gap when no span bridges it.
// ---------------------------------------------------------------------------------
struct AaEdge {          // staged candidate edge, coordinates scaled by 16
  double x0d, g, g63;    // (double) x0, gradient, g * 63.25 (the both-ends-clipped step of polygon.ml:372-379)
  int x0, x1, ymin, ymax;
  int dir, side;         // side: edge_side() against the window the scan is classified in
};
COH_HD AaEdge make_aa_edge(const EdgeRec& e, int wlo, int whi);
COH_HD int pix_of_sub_fast(int n) { const int t = n + 31; return (t + ((t >> 31) & 31)) >> 5; }  // (n + 31) / 32, truncating
struct AaBandLine {      // the crossings of one band line (top or bottom)
  int nl, cl, n, v0, v1; bool hr;
  COH_HD void init() { nl = 0; cl = 0; n = 0; v0 = 0x7FFFFFFF; v1 = 0x7FFFFFFF; hr = false; }
  COH_HD void add(bool pred, int p, int dir, int wlo, int whi) {
    const int px = pix_of_sub_fast(p);
    const bool left = pred && px < wlo, right = pred && px > whi, inside = pred && !(px < wlo) && !(px > whi);
    cl += left ? dir : 0; nl += left ? 1 : 0; hr = hr || right;
    const int nv = inside ? ((p << 1) | (dir > 0)) : 0x7FFFFFFF;
    const int a = imin(v0, nv), b = imax(v0, nv);
    v0 = a; v1 = imin(v1, b);
    n += inside ? 1 : 0;
  }
};
struct AaScan {
  int top, bot, wlo, whi;
  AaBandLine T, B;
  int clo, chi, glo, ghi, nc; bool gap, cplx;
  COH_HD void begin(int y, int wlo_, int whi_) {
    top = 32 * y - 47; bot = top + 63; wlo = wlo_; whi = whi_;
    T.init(); B.init();
    clo = 0x7FFFFFFF; chi = -0x7FFFFFFF; glo = 0; ghi = 0; nc = 0; gap = false; cplx = false;
  }
  COH_HD void edge(const AaEdge& e) {
    const bool horiz = e.ymin == e.ymax;
    const bool cross_top = !horiz && e.ymin < top && e.ymax >= top;      // polygon.ml:338-343 for an edge longer than the band
    const bool cross_bot = !horiz && e.ymax > bot && e.ymin <= bot;
    if (e.side == 1) {
      T.cl += cross_top ? e.dir : 0; T.nl += cross_top ? 1 : 0;
      B.cl += cross_bot ? e.dir : 0; B.nl += cross_bot ? 1 : 0;
      return;
    }
    if (e.side == 2) { T.hr = T.hr || cross_top; B.hr = B.hr || cross_bot; return; }
    const bool active = !(e.ymin > bot || e.ymax < top);
    const int xt = (int)dadd(dadd(e.x0d, dmul(e.g, dadd((double)(top - 1 - e.ymin), 0.25))), 0.5);
    const int xs = cross_top ? xt : e.x0;
    const double step = cross_top ? e.g63 : dmul(e.g, dadd((double)(bot - e.ymin), 0.25));
    const int xb = (int)dadd(dadd((double)xs, step), 0.5);
    T.add(cross_top, xt, e.dir, wlo, whi);
    B.add(cross_bot, xb, e.dir, wlo, whi);
    const int pe = cross_bot ? xb : e.x1;
    const int ca = pix_of_sub_fast(imin(xs, pe) - 16), cb = pix_of_sub_fast(imax(xs, pe) + 16);   // polygon.ml:444-453
    if (active && cb >= wlo && ca <= whi) {
      const bool touch = ca <= chi + 1 && cb >= clo - 1;
      cplx = cplx || (gap && (!touch || (ca <= ghi && cb >= glo)));   // a second gap, or a piece inside the gap
      if (nc > 0 && !touch) {
        gap = true;
        const bool rightof = ca > chi;
        glo = rightof ? chi + 1 : cb + 1; ghi = rightof ? ca - 1 : clo - 1;
      }
      clo = imin(clo, ca); chi = imax(chi, cb); nc++;
    }
  }
  // spans of one band line (polygon.ml:469-512): anything anchored on the window's left / right border, and
  // whether a span reaches across the gap
  COH_HD void line_spans(const AaBandLine& L, int winding, bool& left_any, bool& right_any, bool& bridged) const {
    const int P0 = pix_of_sub_fast(L.v0 >> 1), P1 = pix_of_sub_fast(L.v1 >> 1);
    const bool odd = (L.nl & 1) != 0;
    const int c0 = L.cl + ((L.v0 & 1) ? 1 : -1), c1 = c0 + ((L.v1 & 1) ? 1 : -1);
    const bool e_left = winding == 0 ? (L.cl != 0) : odd;
    const bool e0 = winding == 0 ? (c0 != 0) : !odd;
    const bool e1 = winding == 0 ? (c1 != 0) : odd;
    const bool sL = L.nl > 0 && (L.n > 0 || L.hr) && e_left;    // [wlo, n > 0 ? P0 : whi]
    const bool s0 = L.n >= 1 && e0 && (L.n >= 2 || L.hr);       // [P0, n >= 2 ? P1 : whi]
    const bool s1 = L.n >= 2 && e1 && L.hr;                     // [P1, whi]
    left_any = left_any || sL;
    right_any = right_any || (sL && L.n == 0) || (s0 && L.n == 1) || s1;
    const int bL = L.n > 0 ? P0 : whi, b0 = L.n >= 2 ? P1 : whi;
    bridged = bridged || (sL && bL >= ghi) || (s0 && P0 <= glo && b0 >= ghi) || (s1 && P1 <= glo);
  }
  // true: the row's occupancy inside [wlo, whi] is exactly [lo, hi] minus [glo_, ghi_] (either may be empty: lo > hi)
  COH_HD bool finish(int winding, int& lo, int& hi, int& glo_, int& ghi_) const {
    bool left_any = false, right_any = false, bridged = false;
    line_spans(T, winding, left_any, right_any, bridged);
    line_spans(B, winding, left_any, right_any, bridged);
    glo_ = 1; ghi_ = 0;
    if (cplx || T.n > 2 || B.n > 2) return false;
    if (nc == 0) { lo = (left_any || right_any) ? wlo : 1; hi = (left_any || right_any) ? whi : 0; return true; }
    lo = left_any ? wlo : imax(clo, wlo);
    hi = right_any ? whi : imin(chi, whi);
    if (gap && !bridged) { glo_ = glo; ghi_ = ghi; }
    return true;
  }
};
COH_HD AaEdge make_aa_edge(const EdgeRec& e, int wlo, int whi) {
  AaEdge a;
  a.x0 = e.x0in * 16; a.x1 = e.x1in * 16; a.ymin = e.ymin * 16; a.ymax = e.ymax * 16;
  a.x0d = (double)a.x0; a.g = e.g; a.g63 = dmul(e.g, 63.25);
  a.dir = e.dir; a.side = edge_side(a.x0, a.x1, wlo, whi);
  return a;
}
// table-weighted sum of the run [lo, hi] of super-sampled row j inside the 32 columns starting at w0
COH_HD int aa_interval_sum(const int* prefix_row /*33 ints*/, int lo, int hi, int w0) {
  const int a = imin(imax(lo - w0, 0), 32), b = imin(imax(hi - w0 + 1, 0), 32);
  return b > a ? prefix_row[b] - prefix_row[a] : 0;
}

// ---- constructive planar geometry (render.ml:522-528, 858-981) ----
// shape / minshape words of CPG (op, a, b) from the operands' words; op: 0 Union, 1 Intersection,
// 2 Subtraction, 3 ExclusiveOr
COH_HD void cpg_words(int op, uint32_t SA, uint32_t MA, uint32_t SB, uint32_t MB, uint32_t& S, uint32_t& M) {
  switch (op) {
    case 0: S = SA | SB; M = MA | MB; break;
    case 1: S = SA & SB; M = MA & MB; break;
    case 2: S = SA & ~MB; M = MA & ~SB; break;
    default: S = (SA | SB) & ~(MA & MB); M = (MB & ~SA) | (MA & ~SB); break;
  }
}
// Alpha of a CPG pixel from the operands' matte alphas a, b (255 in the minshape, 0 outside the
// shape).  One expression per operator covers every region the reference splits the sprite into
// (min/min, min/max, max/min, max/max, a only, b only): e.g. Subtraction's "invert b" over
// min_a ∩ max_b is max(0, 255 - b).
COH_HD int cpg_alpha(int op, int a, int b) {
  switch (op) {
    case 0: return a + b > 255 ? 255 : a + b;
    case 1: return a < b ? a : b;
    case 2: return a - b < 0 ? 0 : a - b;
    default: {  // eor (render.ml:858-864)
      const int ia = 255 - a, ib = 255 - b;
      if (a < 128 && b < 128) return a > b ? a : b;
      if (a >= 128 && b < 128) return 255 - (ia > b ? ia : b);
      if (a < 128) return 255 - (a > ib ? a : ib);
      return ia > ib ? ia : ib;
    }
  }
}

// ---------------------------------------------------------------------------------
// Fills (fill.ml:62-140) evaluated per pixel.
// ---------------------------------------------------------------------------------
struct FillRec {
  int kind;            // 0 plain, 1 axial, 2 radial
  uint32_t c0, c1;     // RGBA8 premultiplied
  int flags;           // bit0 ext_s, bit1 ext_e
  double p[6];
};
COH_HD double dsqr(double v) { return dmul(v, v); }
COH_HD_NOINLINE uint32_t fill_lookup(const FillRec& f, int xi, int yi) {
  if (f.kind == 0) return f.c0;
  double x = (double)xi, y = (double)yi;
  if (f.kind == 1) {  // fill.ml:77-93
    double x0 = f.p[0], y0 = f.p[1], x1 = f.p[2], y1 = f.p[3];
    if (x1 == x0 && y1 == y0) return 0u;
    double bottom = dadd(dsqr(dadd(x1, -x0)), dsqr(dadd(y1, -y0)));
    double xp = ddiv(dadd(dmul(dadd(x1, -x0), dadd(x, -x0)), dmul(dadd(y1, -y0), dadd(y, -y0))), bottom);
    if (xp < 0.) return (f.flags & 1) ? f.c0 : 0u;
    if (xp > 1.) return (f.flags & 2) ? f.c1 : 0u;
    return px_dissolve_between(f.c0, f.c1, 255 - (int)dmul(xp, 255.));
  }
  // radial, fill.ml:112-127
#if defined(__CUDA_ARCH__)
#define COH_SQRT(v) __dsqrt_rn(v)
#else
#define COH_SQRT(v) __builtin_sqrt(v)
#endif
  double r = COH_SQRT(dadd(dsqr(dadd(f.p[0], -f.p[2])), dsqr(dadd(f.p[1], -f.p[3]))));
  double r2 = COH_SQRT(dadd(dsqr(dadd(f.p[0], -f.p[4])), dsqr(dadd(f.p[1], -f.p[5]))));
  double diff = dadd(r2, -r);
  double d = COH_SQRT(dadd(dsqr(dadd(f.p[0], -x)), dsqr(dadd(f.p[1], -y))));
  if (d > r2) return (f.flags & 2) ? f.c1 : 0u;
  if (d < r) return (f.flags & 1) ? f.c0 : 0u;
  if (diff == 0.) return f.c0;
  double t = ddiv(dadd(d, -r), diff);
  return px_dissolve_between(f.c0, f.c1, 255 - (int)dmul(t, 255.));
}

}  // namespace coh
