// TEST HARNESS ONLY.  Compiles the product's row arithmetic (raster_core.cuh, the code the
// CUDA kernels execute per lane) with g++ so that the build container — which has no GPU —
// can check it against the oracle on thousands of inputs.  Never part of the product
// library and never a fallback: the C ABI in libcoherence_b200.so has no CPU path.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../coherence_renderer_b200/csrc/raster_core.cuh"

using namespace coh;

static int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

static void runs_to_flat(const std::vector<uint32_t>& bits, int y0, int n_rows, int wx0, int nw, std::vector<int>& flat) {
  for (int r = 0; r < n_rows; r++) {
    std::vector<int> sp;
    int start = 0; bool in = false;
    for (int i = 0; i < nw * 32; i++) {
      bool b = (bits[(size_t)r * nw + (i >> 5)] >> (i & 31)) & 1u;
      if (b && !in) { start = i; in = true; }
      if (!b && in) { sp.push_back(wx0 + start); sp.push_back(i - start); in = false; }
    }
    if (in) { sp.push_back(wx0 + start); sp.push_back(nw * 32 - start); }
    if (!sp.empty()) { flat.push_back(y0 + r); flat.push_back((int)sp.size() / 2); flat.insert(flat.end(), sp.begin(), sp.end()); }
  }
}

extern "C" {
void emul_free(void* p) { std::free(p); }

// shape and minshape of an edge list through scan_row + bit-rows (what k_scan_rows does)
// chunk_words: width (in 32-pixel words) of the windows each scan_row call covers; 0 = the whole row
int emul_shapeminshape(const int32_t* edges, int n, int winding, int chunk_words, int32_t** shp, int64_t* nshp, int32_t** mshp, int64_t* nm) {
  std::vector<EdgeRec> es;
  int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
  for (int i = 0; i < n; i++) {
    es.push_back(make_edge(edges[4 * i], edges[4 * i + 1], edges[4 * i + 2], edges[4 * i + 3]));
    xmin = std::min(xmin, std::min(edges[4 * i], edges[4 * i + 2])); xmax = std::max(xmax, std::max(edges[4 * i], edges[4 * i + 2]));
    ymin = std::min(ymin, std::min(edges[4 * i + 1], edges[4 * i + 3])); ymax = std::max(ymax, std::max(edges[4 * i + 1], edges[4 * i + 3]));
  }
  std::vector<int> fs, fm;
  bool ok = true;
  if (n > 0) {
    int py0 = floordiv(ymin - 16 + 31, 32), py1 = floordiv(ymax + 47, 32);
    int px0 = floordiv(xmin - 16, 32) - 2, px1 = floordiv(xmax + 16 + 31, 32) + 2;
    // two rows / 64 columns of extra margin: the test asserts nothing lands outside the product's box
    int y0 = py0 - 2, n_rows = py1 - py0 + 5;
    int wx0 = floordiv(px0, 32) * 32 - 32, nw = (px1 - wx0) / 32 + 2;
    std::vector<uint32_t> S((size_t)n_rows * nw, 0u), C((size_t)n_rows * nw, 0u);
    int cw = chunk_words > 0 ? chunk_words : nw;
    for (int r = 0; r < n_rows; r++)
      for (int w0 = 0; w0 < nw; w0 += cw) {
        SinkMem sink; sink.wx0 = wx0 + 32 * w0; sink.nwords = std::min(cw, nw - w0); sink.stride = 1;
        sink.S = &S[(size_t)r * nw + w0]; sink.C = &C[(size_t)r * nw + w0];
        ok = scan_row(es.data(), nullptr, n, 1, y0 + r, winding, false, sink.wx0, sink.wx0 + 32 * sink.nwords - 1, sink) && ok;
      }
    // margin must be empty (shape_pixel_box of the product is conservative)
    for (int r = 0; r < n_rows; r++) for (int w = 0; w < nw; w++) {
      bool inside_rows = (y0 + r >= py0 && y0 + r <= py1);
      uint32_t v = S[(size_t)r * nw + w];
      if (!inside_rows && v) return 3;
      for (int b = 0; b < 32; b++) if ((v >> b) & 1u) { int x = wx0 + 32 * w + b; if (x < px0 || x > px1) return 4; }
    }
    std::vector<uint32_t> M(S.size());
    for (size_t i = 0; i < S.size(); i++) M[i] = S[i] & ~C[i];
    runs_to_flat(S, y0, n_rows, wx0, nw, fs);
    runs_to_flat(M, y0, n_rows, wx0, nw, fm);
  }
  *shp = (int32_t*)std::malloc(sizeof(int32_t) * (fs.size() + 1)); std::memcpy(*shp, fs.data(), sizeof(int32_t) * fs.size()); *nshp = (int64_t)fs.size();
  *mshp = (int32_t*)std::malloc(sizeof(int32_t) * (fm.size() + 1)); std::memcpy(*mshp, fm.data(), sizeof(int32_t) * fm.size()); *nm = (int64_t)fm.size();
  return ok ? 0 : 2;
}

static AATable g_aa; static bool g_aa_init = false;
static void init_aa() {
  if (g_aa_init) return;
  int M[32][32];
  for (int x = 1; x <= 32; x++) for (int y = 1; y <= 32; y++) {
    double xp = ((double)(x - 1) * 6.) / 31. - 3., yp = ((double)(y - 1) * 6.) / 31. - 3.;
    M[x - 1][y - 1] = (int)(std::exp(-((xp * xp + yp * yp) / 2.0)) * 255.);
  }
  long total = 0;
  for (int j = 0; j < 32; j++) { g_aa.prefix[j][0] = 0; for (int i = 0; i < 32; i++) { g_aa.prefix[j][i + 1] = g_aa.prefix[j][i] + M[i][j]; total += M[i][j]; } }
  g_aa.volume = (int)((total * 256) / 255);
  g_aa_init = true;
}
// AA opacity of pixels (x0.., y) for a 32-pixel word, the way aa_tile does it (lane j = scaled row j)
int emul_opacity_word(const int32_t* edges, int n, int winding, int x0, int y, uint8_t* out32) {
  init_aa();
  std::vector<EdgeRec> es;
  for (int i = 0; i < n; i++) es.push_back(make_edge(edges[4 * i], edges[4 * i + 1], edges[4 * i + 2], edges[4 * i + 3]));
  uint32_t rows[32][17];
  std::memset(rows, 0, sizeof rows);
  bool ok = true;
  for (int j = 0; j < 32; j++) {
    SinkMem sink; sink.wx0 = 16 * x0 - 32; sink.nwords = 17; sink.stride = 1; sink.S = rows[j]; sink.C = nullptr;
    ok = scan_row(es.data(), nullptr, n, 16, 16 * y - 32 + j, winding, true, sink.wx0, sink.wx0 + 32 * 17 - 1, sink) && ok;
  }
  for (int b = 0; b < 32; b++) {
    int tot = 0;
    for (int j = 0; j < 32; j++) {
      uint32_t lo = rows[j][b >> 1], hi = rows[j][(b >> 1) + 1];
      uint32_t m = (b & 1) ? ((lo >> 16) | (hi << 16)) : lo;
      tot += aa_row_sum(g_aa.prefix[j], m);
    }
    out32[b] = (uint8_t)aa_opacity(tot, g_aa.volume);
  }
  return ok ? 0 : 2;
}
// The register fast path of aa_tile (ScanStateT<3,true> into SinkRow) with its fallback.
// Returns 0 (fast), 1 (fallback used) or -1 on overflow of the general path.
int emul_opacity_word_fast(const int32_t* edges, int n, int winding, int x0, int y, uint8_t* out32) {
  init_aa();
  std::vector<EdgeRec> es;
  for (int i = 0; i < n; i++) es.push_back(make_edge(edges[4 * i], edges[4 * i + 1], edges[4 * i + 2], edges[4 * i + 3]));
  const int wlo = 16 * x0 - 32, whi = wlo + 32 * 17 - 1;
  bool all_fast = true;
  uint32_t rows[32][17];
  std::memset(rows, 0, sizeof rows);
  for (int j = 0; j < 32; j++) {
    ScanStateT<3, true> st; SinkRow sk; sk.wx0 = wlo; sk.nwords = 17; sk.S = rows[j];
    scan_begin(st, 16 * y - 32 + j, true, wlo, whi);
    for (int i = 0; i < n; i++) {
      const EdgeRec& e = es[i];
      int ex0 = e.x0in * 16, ex1 = e.x1in * 16;
      scan_edge(st, ex0, ex1, e.ymin * 16, e.ymax * 16, e.g, e.dir, edge_side(ex0, ex1, wlo, whi), sk);
    }
    all_fast = scan_finish(st, winding, sk) && all_fast;
  }
  if (!all_fast) { int rc = emul_opacity_word(edges, n, winding, x0, y, out32); return rc ? -1 : 1; }
  for (int b = 0; b < 32; b++) {
    int tot = 0;
    for (int j = 0; j < 32; j++) {
      uint32_t lo = rows[j][b >> 1], hi = rows[j][(b >> 1) + 1];
      uint32_t m = (b & 1) ? ((lo >> 16) | (hi << 16)) : lo;
      tot += aa_row_sum(g_aa.prefix[j], m);
    }
    out32[b] = (uint8_t)aa_opacity(tot, g_aa.volume);
  }
  return 0;
}
// The interval form (AaScan): every super-sampled row as one run, classified against the columns under the
// edge pixels `edge` only (as aa_tile narrows its window).  Returns 0 when every row was proven to be one
// interval (out32 written for the pixels of `edge`), 1 when some row was complex (caller falls back), -1 on error.
int emul_opacity_word_interval(const int32_t* edges, int n, int winding, int x0, int y, uint32_t edge, uint8_t* out32) {
  init_aa();
  if (!edge) return 0;
  const int wlo = 16 * x0 - 32;
  const int nlo = wlo + 16 * __builtin_ctz(edge), nhi = wlo + 16 * (31 - __builtin_clz(edge)) + 31;
  std::vector<AaEdge> es;
  for (int i = 0; i < n; i++) es.push_back(make_aa_edge(make_edge(edges[4 * i], edges[4 * i + 1], edges[4 * i + 2], edges[4 * i + 3]), nlo, nhi));
  int lo[32], hi[32], glo[32], ghi[32];
  for (int j = 0; j < 32; j++) {
    AaScan sc; sc.begin(16 * y - 32 + j, nlo, nhi);
    for (int i = 0; i < n; i++) sc.edge(es[i]);
    if (!sc.finish(winding, lo[j], hi[j], glo[j], ghi[j])) return 1;
  }
  for (int b = 0; b < 32; b++) {
    if (!((edge >> b) & 1u)) continue;
    int tot = 0;
    for (int j = 0; j < 32; j++) tot += aa_interval_sum(g_aa.prefix[j], lo[j], hi[j], wlo + 16 * b) - aa_interval_sum(g_aa.prefix[j], glo[j], ghi[j], wlo + 16 * b);
    out32[b] = (uint8_t)aa_opacity(tot, g_aa.volume);
  }
  return 0;
}
// Statistics + self-check over every (row, 32-pixel tile) of one object at a tile grid aligned to multiples of 32
// pixels: the interval form against the general bit-row path for the max-shape pixels (shape - minshape).
// stats[0] += pairs, [1] += complex pairs, [2] += mismatching pixels, [3] += edge pixels
int emul_interval_stats(const int32_t* edges, int n, int winding, int64_t* stats) {
  init_aa();
  std::vector<EdgeRec> es;
  int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
  for (int i = 0; i < n; i++) {
    es.push_back(make_edge(edges[4 * i], edges[4 * i + 1], edges[4 * i + 2], edges[4 * i + 3]));
    xmin = std::min(xmin, std::min(edges[4 * i], edges[4 * i + 2])); xmax = std::max(xmax, std::max(edges[4 * i], edges[4 * i + 2]));
    ymin = std::min(ymin, std::min(edges[4 * i + 1], edges[4 * i + 3])); ymax = std::max(ymax, std::max(edges[4 * i + 1], edges[4 * i + 3]));
  }
  if (n == 0) return 0;
  const int py0 = floordiv(ymin - 16 + 31, 32), py1 = floordiv(ymax + 47, 32);
  const int tx0 = floordiv(floordiv(xmin - 16, 32) - 2, 32), tx1 = floordiv(floordiv(xmax + 16 + 31, 32) + 2, 32);
  for (int y = py0; y <= py1; y++)
    for (int t = tx0; t <= tx1; t++) {
      Sink32 sink; sink.wx0 = 32 * t; sink.S = 0u; sink.C = 0u;
      if (!scan_row(es.data(), nullptr, n, 1, y, winding, false, 32 * t, 32 * t + 31, sink)) return 2;
      const uint32_t edge = sink.C;   // S & ~M with M = S & ~C
      if (!edge) continue;
      stats[0]++; stats[3] += __builtin_popcount(edge);
      uint8_t a[32], b[32];
      int rc = emul_opacity_word_interval(edges, n, winding, 32 * t, y, edge, a);
      if (rc < 0) return 3;
      if (rc == 1) {
        stats[1]++;
        // retry per cluster of edge bits (runs of set bits, split at every zero bit)
        uint32_t m = edge; bool still = false; int ncl = 0;
        while (m) {
          const int s0 = __builtin_ctz(m);
          const uint32_t tz = ~(m >> s0);
          int l = tz ? __builtin_ctz(tz) : 32;
          if (l > 32 - s0) l = 32 - s0;
          const uint32_t cl = (l >= 32 ? 0xFFFFFFFFu : ((1u << l) - 1u)) << s0;
          m &= ~cl; ncl++;
          if (emul_opacity_word_interval(edges, n, winding, 32 * t, y, cl, a) == 1) still = true;
        }
        if (still) stats[4]++;
        stats[5] += ncl;
        continue;
      }
      if (emul_opacity_word(edges, n, winding, 32 * t, y, b)) return 4;
      for (int k = 0; k < 32; k++) if (((edge >> k) & 1u) && a[k] != b[k]) stats[2]++;
    }
  return 0;
}
uint32_t emul_over(uint32_t a, uint32_t b) { return px_over(a, b); }
uint32_t emul_dissolve(uint32_t c, int d) { return px_dissolve(c, d); }
uint32_t emul_dissolve_between(uint32_t a, uint32_t b, int alpha) { return px_dissolve_between(a, b, alpha); }
uint32_t emul_alpha_over(uint32_t a, uint32_t b) { return alpha_over(a, b); }
uint32_t emul_fill(int kind, uint32_t c0, uint32_t c1, int flags, const double* p, int x, int y) {
  FillRec f; f.kind = kind; f.c0 = c0; f.c1 = c1; f.flags = flags; for (int i = 0; i < 6; i++) f.p[i] = p[i];
  return fill_lookup(f, x, y);
}
}
