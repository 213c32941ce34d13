// coherence_b200.cu — C ABI (include/coherence_b200.h) over the sm_100a kernels.
// Host glue only: buffer management, launches, error reporting.  No CPU fallback: every
// compute entry point needs a live CUDA context and fails loudly otherwise.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include "../../include/coherence_b200.h"
#include "kernels.cuh"

using namespace coh;

// ---------------------------------------------------------------------------------------
struct DevShape {
  int y0 = 0, n_rows = 0;       // rows y0 .. y0+n_rows-1 (row lists may be empty)
  int* row_ptr = nullptr;       // device, n_rows+1
  int2* spans = nullptr;        // device
  int n_spans = 0;
  long long card = 0;           // pixels
  int bx0 = 0, by0 = 0, bx1 = -1, by1 = -1;  // tight bounds (valid if n_spans > 0)
};
struct DevScene {
  int n_objs = 0, n_leaves = 0, n_edges = 0, n_points = 0;
  ObjRec* objs = nullptr;
  int* leaves = nullptr;
  int4* leaf_box = nullptr;     // conservative device-space pixel box per leaf (binning reads these, coalesced)
  EdgeRec* edges = nullptr;
  int2* points = nullptr;
  uint8_t* stamps = nullptr;
  int* rowedge_ptr = nullptr;   // K1 edge binning (CSR over (path object, pixel row))
  int* rowedge_idx = nullptr;
  int2* brush_ranges = nullptr; // per (stroke, cell of its box): [first, last] stamp index reaching the cell
  uint32_t* conv_bits = nullptr; // Convolved objects: shape / minshape bit-rows
  uint32_t* conv_px = nullptr;   // Convolved objects: pre-convolved canvases
  std::vector<ObjRec> h_objs;
  std::vector<int64_t> ids;      // cache key (Id.idset) of every record
  std::vector<int> rec_of_abi;   // record index of every object of the ABI array (-1: GROUP_END / dropped)
  std::vector<int> group_last;   // for group records: last record index inside the group
  // Filters (render.ml:37-48): top-level members of the scene list that are not leaves.  `pos` = number of
  // ordinary scene leaves in front of the filter; the leaves are ordered [scene | reading scenes | background].
  struct FilterRec { int pos, kind, kernel_kind, r, first, count, winding; uint32_t colour; int read0, read1; int bx0, by0, bx1, by1; int abi; };
  std::vector<FilterRec> filters;
  // Group shapes (render.ml:476-496 caches them under the group's id): kept per scene, in the frame the group
  // had when the entry was made; moving the whole group only changes the offset applied on the way out.
  struct GroupShape { DevShape* shape; int offx, offy; };
  std::map<int, GroupShape> group_shape;
  std::vector<int2> group_off;   // per record: translation applied to the whole group since scene creation
  int n_scene_leaves = 0;        // ordinary leaves of the scene list
  int n_front_leaves = 0;        // + leaves of reading-scene groups (the background list follows)
  std::vector<int> h_leaves;
  bool has_fancy = false;    // some object has a gradient / radial fill
  int extras = 0;            // walker variant: 0 polygons / primitives, 1 + brush / Convolved, 2 + CPG / filters
  size_t items_total = 0, coarse_total = 0; bool coarse_total_valid = false;
  int items_for_W = -1, items_for_H = -1, items_for_y0 = -1, items_for_y1 = -1;
};

// Cache (cache.ml:57-83): entries keyed by object id hold device-resident span sets
// (shape, minshape); an alias entry refers to another id with an integer translation.
struct CacheEntry {
  bool alias = false; int dx = 0, dy = 0; int64_t target = 0;
  DevShape* shape = nullptr; DevShape* minshape = nullptr; bool has = false;
  size_t bytes = 0; uint64_t lastused = 0;
};

struct coh_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  AATable* d_aa = nullptr;
  int* d_error = nullptr;
  int* h_error = nullptr;  // pinned
  // framebuffer
  Frame fr{0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t* fb = nullptr;
  uint32_t* u_out = nullptr;   // bit-frame of `u` after the scene pass
  uint32_t* u_init = nullptr;  // bit-frame of an arbitrary update shape
  bool use_u_init = false;
  bool have_u = false;
  // binning scratch
  int n_cells_cap = 0;
  int2* cell_head = nullptr;  // per cell: colour + flags when the cell is one opaque covering primitive
  int2* cell_rng = nullptr;   // per cell [start, end) into cell_items
  // large scenes: coarse level of the two-level binning (leaf positions per coarse cell)
  int* coarse_items = nullptr; int* coarse_counts = nullptr; int* coarse_off = nullptr; size_t coarse_cap = 0, coarse_cells_cap = 0;
  uint32_t* peer_fb[COH_MAX_PEERS] = {nullptr}; int n_peers = 0;  // coh_fb_set_peers
  // three-phase frames: per (cell item, row) pair
  uint2* pre_sc = nullptr; int4* pre_list = nullptr; int* pre_n = nullptr; uint8_t* pre_op = nullptr; size_t pre_cap = 0;
  // asynchronous read-back (coh_fb_read_rgba_async): two staging buffers, a copy stream
  cudaStream_t copy_stream = nullptr;
  uint32_t* stage[2] = {nullptr, nullptr}; size_t stage_cap[2] = {0, 0}; bool stage_busy[2] = {false, false};
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  int stage_next = 0;
  int* cell_items = nullptr; size_t cell_items_cap = 0;
  int* item_cell = nullptr;   // cell of every list entry (small-scene binning only)
  int* h_total = nullptr;  // pinned
  // cross-tile carry for fancy fills
  int* queue = nullptr; int* order_hist = nullptr; int* cell_order = nullptr; int n_sms = 0;
  int* carry_done = nullptr; int* carry_cnt = nullptr; int2* carry_ent = nullptr;
  size_t carry_slots = 0; int epoch = 0;
  bool own_stream = true, own_fb = true;
  // coherence cache (HBM-resident span sets)
  std::map<int64_t, CacheEntry> cache;
  bool usecache = true; size_t cache_max = 50u * 1024u * 1024u, cache_size = 0; uint64_t cache_timer = 0;  // cache.ml:72-73
  int64_t shphit = 0, shpmis = 0;
  // optional per-kernel timing (CUDA events on the launching stream)
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // bin start, walk start, walk end, spare
  double walk_ms_sum = 0, bin_ms_sum = 0; long timed_frames = 0;
  bool ev_pending = false;
};

static std::string g_init_err;

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char b_[512];                                                                           \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = b_;                                                                          \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)
#define FAIL(msg) do { ctx->err = (msg); return 1; } while (0)
// Device memory comes from the stream-ordered pool (cudaMallocAsync): allocation and release are
// ordered on the context's stream and cost microseconds instead of a device-wide synchronisation.
#define DMALLOC(ptr, bytes) cudaMallocAsync((void**)(ptr), (bytes), ctx->stream)
#define DFREE(ptr) do { if (ptr) cudaFreeAsync((void*)(ptr), ctx->stream); } while (0)
#define LAUNCHED() do { ctx->launches++; CK(cudaGetLastError()); } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
// Exclusive scan of n ints into out[0..n] on the context's stream.  Up to 64 Ki elements every block
// reduces its own prefix (one launch); beyond that the 1024-element tile sums are scanned recursively.
static int exclusive_scan(coh_ctx* ctx, const int* in, int* out, int n, int* hist);
static inline int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

static int exclusive_scan(coh_ctx* ctx, const int* in, int* out, int n, int* hist) {
  const int tiles = std::max(1, cdiv(n, 1024));
  if (n <= 65536) {
    k_exclusive_scan<<<tiles, 1024, 0, ctx->stream>>>(in, out, n, hist, nullptr); LAUNCHED();
    return 0;
  }
  int *sums = nullptr, *offs = nullptr;
  CK(DMALLOC(&sums, sizeof(int) * tiles)); CK(DMALLOC(&offs, sizeof(int) * (tiles + 1)));
  k_tile_sums<<<tiles, 1024, 0, ctx->stream>>>(in, sums, n); LAUNCHED();
  if (exclusive_scan(ctx, sums, offs, tiles, nullptr)) return 1;
  k_exclusive_scan<<<tiles, 1024, 0, ctx->stream>>>(in, out, n, hist, offs); LAUNCHED();
  DFREE(sums); DFREE(offs);
  return 0;
}

// AA table (polygon.ml:616-651): maintable via exp on the host once; prefix sums per scaled row.
static void build_aa_table(AATable& t) {
  int M[32][32];
  for (int x = 1; x <= 32; x++)
    for (int y = 1; y <= 32; y++) {
      double xp = ((double)(x - 1) * 6.) / 31. - 3., yp = ((double)(y - 1) * 6.) / 31. - 3.;
      M[x - 1][y - 1] = (int)(exp(-((xp * xp + yp * yp) / 2.0)) * 255.);
    }
  long total = 0;
  for (int j = 0; j < 32; j++) {
    t.prefix[j][0] = 0;
    for (int i = 0; i < 32; i++) { t.prefix[j][i + 1] = t.prefix[j][i] + M[i][j]; total += M[i][j]; }
  }
  t.volume = (int)((total * 256) / 255);
  if (t.volume != AA_VOLUME) { fprintf(stderr, "coherence_b200: AA table volume %d != %d\n", t.volume, AA_VOLUME); abort(); }
}

extern "C" {

const char* coh_last_error(coh_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_err.c_str(); }

int coh_init(int device, coh_ctx** out) {
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_init_err = std::string("coh_init: no CUDA device available (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return 1;
  }
  coh_ctx* ctx = new coh_ctx();
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  ctx->device = device;
  auto bail = [&](const char* what, cudaError_t err) { g_init_err = std::string(what) + ": " + cudaGetErrorString(err); delete ctx; return 1; };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;  // keep freed blocks in the pool instead of returning them to the driver
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  AATable t; build_aa_table(t);
  if ((e = DMALLOC(&ctx->d_aa, sizeof(AATable))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemcpyAsync(ctx->d_aa, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail("cudaMemcpy", e);
  if ((e = DMALLOC(&ctx->d_error, sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
  cudaMemsetAsync(ctx->d_error, 0, sizeof(int), ctx->stream);
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail("cudaStreamSynchronize", e);
  if ((e = cudaMallocHost(&ctx->h_error, sizeof(int))) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = cudaMallocHost(&ctx->h_total, sizeof(int))) != cudaSuccess) return bail("cudaMallocHost", e);
  *out = ctx;
  return 0;
}

int coh_cache_clear(coh_ctx* ctx);
int coh_shutdown(coh_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  coh_fb_read_wait(ctx);
  DFREE(ctx->stage[0]); DFREE(ctx->stage[1]);
  if (ctx->copy_stream) {
    cudaStreamDestroy(ctx->copy_stream);
    for (int k = 0; k < 2; k++) { cudaEventDestroy(ctx->ev_ready[k]); cudaEventDestroy(ctx->ev_done[k]); }
  }
  coh_cache_clear(ctx);
  DFREE(ctx->d_aa); DFREE(ctx->d_error); cudaFreeHost(ctx->h_error); cudaFreeHost(ctx->h_total);
  if (ctx->own_fb) DFREE(ctx->fb);
  DFREE(ctx->u_out); DFREE(ctx->u_init);
  DFREE(ctx->cell_items); DFREE(ctx->cell_head); DFREE(ctx->cell_rng);
  DFREE(ctx->coarse_items); DFREE(ctx->coarse_counts); DFREE(ctx->coarse_off);
  DFREE(ctx->pre_sc); DFREE(ctx->pre_list); DFREE(ctx->pre_n); DFREE(ctx->pre_op); DFREE(ctx->item_cell);
  DFREE(ctx->order_hist); DFREE(ctx->cell_order); DFREE(ctx->carry_done); DFREE(ctx->carry_cnt); DFREE(ctx->carry_ent);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  delete ctx;
  return 0;
}
int coh_device_name(coh_ctx* ctx, char* buf, int cap) {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, ctx->device));
  snprintf(buf, cap, "%s (sm_%d%d, %d SMs)", p.name, p.major, p.minor, p.multiProcessorCount);
  return 0;
}
void* coh_stream(coh_ctx* ctx) { return (void*)ctx->stream; }
int coh_set_stream(coh_ctx* ctx, void* stream) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)stream; ctx->own_stream = false;
  return 0;
}
static int drain_timing(coh_ctx* ctx) {
  if (!ctx->ev_pending) return 0;
  CK(cudaEventSynchronize(ctx->ev[2]));
  float a = 0, b = 0;
  CK(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
  CK(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
  ctx->bin_ms_sum += a; ctx->walk_ms_sum += b; ctx->timed_frames++;
  ctx->ev_pending = false;
  return 0;
}
int coh_set_timing(coh_ctx* ctx, int32_t on) {
  CK(cudaSetDevice(ctx->device));
  if (drain_timing(ctx)) return 1;
  if (on && !ctx->ev[0]) for (int i = 0; i < 4; i++) CK(cudaEventCreate(&ctx->ev[i]));
  ctx->timing = on != 0; ctx->walk_ms_sum = 0; ctx->bin_ms_sum = 0; ctx->timed_frames = 0;
  return 0;
}
int coh_get_timing(coh_ctx* ctx, double* walk_ms_avg, double* bin_ms_avg, int64_t* frames) {
  CK(cudaSetDevice(ctx->device));
  if (drain_timing(ctx)) return 1;
  *frames = ctx->timed_frames;
  *walk_ms_avg = ctx->timed_frames ? ctx->walk_ms_sum / ctx->timed_frames : 0.;
  *bin_ms_avg = ctx->timed_frames ? ctx->bin_ms_sum / ctx->timed_frames : 0.;
  return 0;
}
int64_t coh_launch_count(coh_ctx* ctx) { return ctx->launches; }
static int check_error_flag(coh_ctx* ctx, const char* what);
int coh_mem_in_use(coh_ctx* ctx, int64_t* bytes) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
  uint64_t used = 0;
  CK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
  *bytes = (int64_t)used;
  return 0;
}
int coh_sync(coh_ctx* ctx) {
  CK(cudaSetDevice(ctx->device));
  return check_error_flag(ctx, "coh_sync");  // synchronises the stream and reports deferred kernel-side failures
}

// ---- colour codec: colour.ml:99-172 (host-side pure functions for the OCaml boundary) ----
int32_t coh_colour_of_rgba8(uint32_t w) {
  int r8 = w & 255, g8 = (w >> 8) & 255, b8 = (w >> 16) & 255, a8 = w >> 24;
  int r = r8 >> 1, g = g8 >> 1, b = b8 >> 1, a = a8 >> 1;
  int rl = r8 & 1, gl = g8 & 1, bl = b8 & 1, al = a8 & 1;
  auto cat = [](int p, int q, int s, int t) { return (p << 21) | (q << 14) | (s << 7) | t; };
  if (r != a && g != a && b != a)
    return (rl << 29) | (gl << 28) | (bl ? (al ? cat(r, g, b, a) : cat(r, g, a, b)) : (al ? cat(r, a, b, g) : cat(a, g, b, r)));
  int tail = r == a ? cat(0, g, b, a) : g == a ? cat(0, r, b, a) : cat(0, r, g, a);
  return (1 << 30) | (rl << 29) | (gl << 28) | (bl << 27) | (al << 26) | ((r == a) << 25) | ((g == a) << 24) | ((b == a) << 23) | tail;
}
uint32_t coh_rgba8_of_colour(int32_t c) {
  int r = 0, g = 0, b = 0, a = 0, rl = (c >> 29) & 1, gl = (c >> 28) & 1, bl = 0, al = 0;
  int c3 = (c >> 21) & 127, c2 = (c >> 14) & 127, c1 = (c >> 7) & 127, c0 = c & 127;
  if (!(c & (1 << 30))) {
    int m;  // index of the maximum, colour.ml:86-96
    if (c3 > c2) m = (c1 > c0) ? (c3 > c1 ? 0 : 2) : (c3 > c0 ? 0 : 3);
    else m = (c1 > c0) ? (c2 > c1 ? 1 : 2) : (c2 > c0 ? 1 : 3);
    switch (m) {
      case 3: bl = 1; al = 1; r = c3; g = c2; b = c1; a = c0; break;
      case 2: bl = 1; al = 0; r = c3; g = c2; a = c1; b = c0; break;
      case 1: bl = 0; al = 1; r = c3; a = c2; b = c1; g = c0; break;
      default: bl = 0; al = 0; a = c3; g = c2; b = c1; r = c0; break;
    }
  } else {
    bl = (c >> 27) & 1; al = (c >> 26) & 1; a = c0;
    if (c & (1 << 25)) { r = a; g = c2; b = c1; }
    else if (c & (1 << 24)) { g = a; r = c2; b = c1; }
    else { b = a; r = c2; g = c1; }
  }
  return (uint32_t)((r << 1) | rl) | ((uint32_t)((g << 1) | gl) << 8) | ((uint32_t)((b << 1) | bl) << 16) | ((uint32_t)((a << 1) | al) << 24);
}

// ---------------------------------------------------------------------------------------
// Shapes
// ---------------------------------------------------------------------------------------
static void free_shape(coh_ctx* ctx, DevShape* s) {
  if (!s) return;
  DFREE(s->row_ptr); DFREE(s->spans);
  delete s;
}
int coh_shape_free(coh_ctx* ctx, coh_shape_t h) {
  CK(cudaSetDevice(ctx->device));
  free_shape(ctx, (DevShape*)h);
  return 0;
}

// Bit-frame [n_rows][nw] (device) -> span set.  Consumes nothing; returns 0 handle for the empty set.
static int shape_from_bits(coh_ctx* ctx, const uint32_t* bits, int y0, int n_rows, int wx0, int nw, coh_shape_t* out) {
  *out = 0;
  if (n_rows <= 0 || nw <= 0) return 0;
  int* counts = nullptr; int* ptr = nullptr; unsigned long long* d_card = nullptr;
  CK(DMALLOC(&counts, sizeof(int) * n_rows));
  CK(DMALLOC(&ptr, sizeof(int) * (n_rows + 1)));
  CK(DMALLOC(&d_card, sizeof(unsigned long long)));
  CK(cudaMemsetAsync(d_card, 0, sizeof(unsigned long long), ctx->stream));
  k_count_runs<<<cdiv(n_rows, 128), 128, 0, ctx->stream>>>(bits, n_rows, nw, counts, d_card); LAUNCHED();
  if (exclusive_scan(ctx, counts, ptr, n_rows, nullptr)) return 1;
  std::vector<int> h_ptr(n_rows + 1);
  unsigned long long card = 0;
  CK(cudaMemcpyAsync(h_ptr.data(), ptr, sizeof(int) * (n_rows + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&card, d_card, sizeof card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(counts); DFREE(d_card);
  int total = h_ptr[n_rows];
  if (total == 0) { DFREE(ptr); return 0; }
  // trim empty rows at both ends so that y0 / n_rows are tight
  int first = 0, last = n_rows - 1;
  while (h_ptr[first + 1] == h_ptr[first]) first++;
  while (h_ptr[last + 1] == h_ptr[last]) last--;
  DevShape* s = new DevShape();
  s->n_spans = total; s->card = (long long)card;
  CK(DMALLOC(&s->spans, sizeof(int2) * total));
  k_fill_runs<<<cdiv(n_rows, 128), 128, 0, ctx->stream>>>(bits, n_rows, nw, wx0, ptr, s->spans); LAUNCHED();
  s->y0 = y0 + first; s->n_rows = last - first + 1;
  CK(DMALLOC(&s->row_ptr, sizeof(int) * (s->n_rows + 1)));
  CK(cudaMemcpyAsync(s->row_ptr, ptr + first, sizeof(int) * (s->n_rows + 1), cudaMemcpyDeviceToDevice, ctx->stream));
  // bounds: x extremes need the spans; take them from a host copy (export path, not hot)
  std::vector<int2> h_spans(total);
  CK(cudaMemcpyAsync(h_spans.data(), s->spans, sizeof(int2) * total, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(ptr);
  s->by0 = s->y0; s->by1 = s->y0 + s->n_rows - 1; s->bx0 = INT32_MAX; s->bx1 = INT32_MIN;
  for (int r = first; r <= last; r++) {
    if (h_ptr[r + 1] > h_ptr[r]) {
      s->bx0 = std::min(s->bx0, h_spans[h_ptr[r]].x);
      const int2& l = h_spans[h_ptr[r + 1] - 1];
      s->bx1 = std::max(s->bx1, l.x + l.y - 1);
    }
  }
  *out = (coh_shape_t)s;
  return 0;
}
// span set -> freshly allocated bit-frame covering rows [y0, y0+n_rows) and words from pixel wx0
static int bits_from_shape(coh_ctx* ctx, const DevShape* s, int y0, int n_rows, int wx0, int nw, uint32_t** out) {
  uint32_t* bits = nullptr;
  CK(DMALLOC(&bits, sizeof(uint32_t) * (size_t)n_rows * nw));
  CK(cudaMemsetAsync(bits, 0, sizeof(uint32_t) * (size_t)n_rows * nw, ctx->stream));
  if (s && s->n_spans > 0) {
    k_spans_to_bits<<<cdiv(n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, s->y0, s->n_rows, y0, n_rows, wx0, nw, bits);
    LAUNCHED();
  }
  *out = bits;
  return 0;
}

int coh_shape_box(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (w == 0 && h == 0) return 0;                       // sprite.ml:463
  if (w < 0 || h < 0) FAIL("Sprite.box: negative argument.");  // sprite.ml:464
  if (w == 0 || h == 0) return 0;
  std::vector<int> flat;
  for (int r = 0; r < h; r++) { flat.push_back(y + r); flat.push_back(1); flat.push_back(x); flat.push_back(w); }
  return coh_shape_import(ctx, flat.data(), (int64_t)flat.size(), out);
}
int coh_shape_import(coh_ctx* ctx, const int32_t* flat, int64_t n, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (n == 0) return 0;
  // validate canonical form (sprite.ml:201-239) while building the CSR
  std::vector<int> ys; std::vector<int> cnt; std::vector<int2> spans;
  int64_t i = 0; long long card = 0;
  int bx0 = INT32_MAX, bx1 = INT32_MIN;
  while (i < n) {
    if (i + 2 > n) FAIL("shape import: truncated row header");
    int y = flat[i], k = flat[i + 1]; i += 2;
    if (k <= 0) FAIL("shape import: malformed shape (empty spanline)");
    if (!ys.empty() && y <= ys.back()) FAIL("shape import: malformed shape (rows not increasing)");
    if (i + 2 * (int64_t)k > n) FAIL("shape import: truncated spans");
    for (int q = 0; q < k; q++, i += 2) {
      int x = flat[i], l = flat[i + 1];
      if (l <= 0) FAIL("shape import: malformed shape (span length)");
      if (q && x <= spans.back().x + spans.back().y) FAIL("shape import: malformed shape (spans overlap or abut)");
      spans.push_back(make_int2(x, l)); card += l;
      bx0 = std::min(bx0, x); bx1 = std::max(bx1, x + l - 1);
    }
    ys.push_back(y); cnt.push_back(k);
  }
  DevShape* s = new DevShape();
  s->y0 = ys.front(); s->n_rows = ys.back() - ys.front() + 1;
  std::vector<int> ptr(s->n_rows + 1, 0);
  for (size_t r = 0; r < ys.size(); r++) ptr[ys[r] - s->y0 + 1] = cnt[r];
  for (int r = 0; r < s->n_rows; r++) ptr[r + 1] += ptr[r];
  s->n_spans = (int)spans.size(); s->card = card;
  s->bx0 = bx0; s->bx1 = bx1; s->by0 = ys.front(); s->by1 = ys.back();
  CK(DMALLOC(&s->row_ptr, sizeof(int) * ptr.size()));
  CK(DMALLOC(&s->spans, sizeof(int2) * spans.size()));
  CK(cudaMemcpyAsync(s->row_ptr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(s->spans, spans.data(), sizeof(int2) * spans.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *out = (coh_shape_t)s;
  return 0;
}
static int download_shape(coh_ctx* ctx, const DevShape* s, std::vector<int>& ptr, std::vector<int2>& spans) {
  ptr.resize(s->n_rows + 1); spans.resize(s->n_spans);
  CK(cudaMemcpyAsync(ptr.data(), s->row_ptr, sizeof(int) * ptr.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(spans.data(), s->spans, sizeof(int2) * spans.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int coh_shape_export_size(coh_ctx* ctx, coh_shape_t h, int64_t* n) {
  CK(cudaSetDevice(ctx->device));
  *n = 0;
  if (!h) return 0;
  DevShape* s = (DevShape*)h;
  std::vector<int> ptr; std::vector<int2> spans;
  if (download_shape(ctx, s, ptr, spans)) return 1;
  int64_t rows = 0;
  for (int r = 0; r < s->n_rows; r++) rows += ptr[r + 1] > ptr[r];
  *n = 2 * rows + 2 * (int64_t)s->n_spans;
  return 0;
}
int coh_shape_export(coh_ctx* ctx, coh_shape_t h, int32_t* flat, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  if (!h) return 0;
  DevShape* s = (DevShape*)h;
  std::vector<int> ptr; std::vector<int2> spans;
  if (download_shape(ctx, s, ptr, spans)) return 1;
  int64_t k = 0;
  for (int r = 0; r < s->n_rows; r++) {
    int c = ptr[r + 1] - ptr[r];
    if (!c) continue;
    if (k + 2 + 2 * c > cap) FAIL("coh_shape_export: buffer too small");
    flat[k++] = s->y0 + r; flat[k++] = c;
    for (int q = ptr[r]; q < ptr[r + 1]; q++) { flat[k++] = spans[q].x; flat[k++] = spans[q].y; }
  }
  *n_out = k;
  return 0;
}
int coh_shape_bounds(coh_ctx* ctx, coh_shape_t h, int32_t box[4], int32_t* is_null) {
  (void)ctx;
  DevShape* s = (DevShape*)h;
  *is_null = !s;
  if (s) { box[0] = s->bx0; box[1] = s->by0; box[2] = s->bx1; box[3] = s->by1; }
  return 0;
}
int coh_shape_card(coh_ctx* ctx, coh_shape_t h, int64_t* n) {
  (void)ctx;
  *n = h ? ((DevShape*)h)->card : 0;
  return 0;
}

// Binary set algebra through bit-frames over the union bounding box (K3).
static int shape_binop(coh_ctx* ctx, coh_shape_t ha, coh_shape_t hb, int op, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  DevShape* a = (DevShape*)ha; DevShape* b = (DevShape*)hb;
  if (!a && !b) return 0;
  if (!a && op != 0) return 0;       // {} - b = {} ; {} & b = {}
  if (!b && op == 2) return 0;
  int x0 = INT32_MAX, x1 = INT32_MIN, y0 = INT32_MAX, y1 = INT32_MIN;
  for (DevShape* s : {a, b}) if (s) { x0 = std::min(x0, s->bx0); x1 = std::max(x1, s->bx1); y0 = std::min(y0, s->by0); y1 = std::max(y1, s->by1); }
  int wx0 = floordiv(x0, 32) * 32, nw = (x1 - wx0) / 32 + 1, n_rows = y1 - y0 + 1;
  uint32_t *ba = nullptr, *bb = nullptr;
  if (bits_from_shape(ctx, a, y0, n_rows, wx0, nw, &ba)) return 1;
  if (bits_from_shape(ctx, b, y0, n_rows, wx0, nw, &bb)) return 1;
  size_t n = (size_t)n_rows * nw;
  k_bitop<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ba, bb, ba, n, op); LAUNCHED();
  int rc = shape_from_bits(ctx, ba, y0, n_rows, wx0, nw, out);
  DFREE(ba); DFREE(bb);
  return rc;
}
int coh_shape_union(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out) { return shape_binop(ctx, a, b, 0, out); }
int coh_shape_difference(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out) { return shape_binop(ctx, a, b, 1, out); }
int coh_shape_intersection(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out) { return shape_binop(ctx, a, b, 2, out); }

int coh_shape_translate(coh_ctx* ctx, coh_shape_t h, int32_t dx, int32_t dy, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (!h) return 0;
  DevShape* s = (DevShape*)h;
  DevShape* t = new DevShape(*s);
  t->y0 += dy; t->bx0 += dx; t->bx1 += dx; t->by0 += dy; t->by1 += dy;
  CK(DMALLOC(&t->row_ptr, sizeof(int) * (s->n_rows + 1)));
  CK(DMALLOC(&t->spans, sizeof(int2) * s->n_spans));
  CK(cudaMemcpyAsync(t->row_ptr, s->row_ptr, sizeof(int) * (s->n_rows + 1), cudaMemcpyDeviceToDevice, ctx->stream));
  k_translate_spans<<<cdiv(s->n_spans, 256), 256, 0, ctx->stream>>>(s->spans, t->spans, s->n_spans, dx); LAUNCHED();
  *out = (coh_shape_t)t;
  return 0;
}
static int bloat_impl(coh_ctx* ctx, const DevShape* s, int x0, int y0, int x1, int y1, int m, int n, bool complement_in_box,
                      coh_shape_t* out) {
  // frame = box [x0..x1] x [y0..y1] grown by (m, n) on every side
  int fx0 = x0 - m, fy0 = y0 - n, fx1 = x1 + m, fy1 = y1 + n;
  int wx0 = floordiv(fx0, 32) * 32, nw = (fx1 - wx0) / 32 + 1, n_rows = fy1 - fy0 + 1;
  uint32_t *in = nullptr, *tmp = nullptr;
  if (bits_from_shape(ctx, s, fy0, n_rows, wx0, nw, &in)) return 1;
  CK(DMALLOC(&tmp, sizeof(uint32_t) * (size_t)n_rows * nw));
  size_t nwords = (size_t)n_rows * nw;
  if (complement_in_box) {
    // erode (sprite.ml:1867-1877): inverse = enclosing - shp, bloated, then shp - bloated
    uint32_t* box = nullptr;
    CK(DMALLOC(&box, sizeof(uint32_t) * nwords));
    dim3 g(cdiv(nw, 128), n_rows);
    k_fill_box_bits<<<g, 128, 0, ctx->stream>>>(box, n_rows, nw, wx0, fy0, fx0, fy0, fx1, fy1); LAUNCHED();
    k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(box, in, box, nwords, 1); LAUNCHED();  // inverse
    k_dilate<<<g, 128, 0, ctx->stream>>>(box, tmp, n_rows, nw, m, n); LAUNCHED();
    k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(in, tmp, tmp, nwords, 1); LAUNCHED();    // shp - bloated
    DFREE(box);
  } else {
    dim3 g(cdiv(nw, 128), n_rows);
    k_dilate<<<g, 128, 0, ctx->stream>>>(in, tmp, n_rows, nw, m, n); LAUNCHED();
  }
  int rc = shape_from_bits(ctx, tmp, fy0, n_rows, wx0, nw, out);
  DFREE(in); DFREE(tmp);
  return rc;
}
int coh_shape_bloat(coh_ctx* ctx, coh_shape_t h, int32_t m, int32_t n, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (!h) return 0;
  if (m < 0 || n < 0) FAIL("Sprite.bloat: negative radius");
  DevShape* s = (DevShape*)h;
  return bloat_impl(ctx, s, s->bx0, s->by0, s->bx1, s->by1, m, n, false, out);
}
int coh_shape_erode(coh_ctx* ctx, coh_shape_t h, int32_t m, int32_t n, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (!h) return 0;
  if (m < 0 || n < 0) FAIL("Sprite.erode: negative radius");
  DevShape* s = (DevShape*)h;
  return bloat_impl(ctx, s, s->bx0, s->by0, s->bx1, s->by1, m, n, true, out);
}

// ---------------------------------------------------------------------------------------
// Polygon
// ---------------------------------------------------------------------------------------
struct EdgeBox { int xmin, xmax, ymin, ymax; };
static EdgeBox edge_bounds(const int32_t* e, int n) {
  EdgeBox b{INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
  for (int i = 0; i < n; i++) {
    b.xmin = std::min(b.xmin, std::min(e[4 * i], e[4 * i + 2])); b.xmax = std::max(b.xmax, std::max(e[4 * i], e[4 * i + 2]));
    b.ymin = std::min(b.ymin, std::min(e[4 * i + 1], e[4 * i + 3])); b.ymax = std::max(b.ymax, std::max(e[4 * i + 1], e[4 * i + 3]));
  }
  return b;
}
// Conservative pixel box of the shape of an edge list: a row y is touched iff its band
// [32y-47, 32y+16] meets [ymin, ymax]; columns from the widened coverage (polygon.ml:444-453)
// plus two pixels of slack: band crossings are rounded by truncation toward zero and the
// bottom crossing of a doubly clipped edge restarts from the rounded top crossing
// (polygon.ml:365-379), so a crossing can leave the edge's x range by up to 3 sub-bins, and
// pix_of_sub itself truncates toward zero on negative sub-bins.
static void shape_pixel_box(const EdgeBox& b, int& px0, int& py0, int& px1, int& py1) {
  py0 = floordiv(b.ymin - 16 + 31, 32);   // smallest y with 32y+16 >= ymin
  py1 = floordiv(b.ymax + 47, 32);        // largest y with 32y-47 <= ymax
  px0 = floordiv(b.xmin - 16, 32) - 2;
  px1 = floordiv(b.xmax + 16 + 31, 32) + 2;
}
static int upload_edges(coh_ctx* ctx, const int32_t* edges, int n, EdgeRec** out) {
  int4* raw = nullptr;
  CK(DMALLOC(&raw, sizeof(int4) * std::max(n, 1)));
  CK(DMALLOC(out, sizeof(EdgeRec) * std::max(n, 1)));
  if (n > 0) {
    CK(cudaMemcpyAsync(raw, edges, sizeof(int4) * n, cudaMemcpyHostToDevice, ctx->stream));
    k_prep_edges<<<cdiv(n, 256), 256, 0, ctx->stream>>>(raw, *out, n); LAUNCHED();
  }
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(raw);
  return 0;
}
static int check_error_flag(coh_ctx* ctx, const char* what) {
  CK(cudaMemcpyAsync(ctx->h_error, ctx->d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (*ctx->h_error) {
    cudaMemsetAsync(ctx->d_error, 0, sizeof(int), ctx->stream);
    ctx->err = std::string(what) + ": an object has more than " + std::to_string(COH_MAXX) + " band crossings inside one tile window of a row (COH_MAXX), or more than " + std::to_string(CARRY_CAP) + " fancy-fill edge runs cross one tile border (CARRY_CAP)";
    return 1;
  }
  return 0;
}

// scan-convert device-resident prepared edges inside a pixel box into (shape, minshape) span sets
static int shapes_from_device_edges(coh_ctx* ctx, const EdgeRec* d_edges, int n_edges, int winding, int px0, int py0,
                                    int px1, int py1, coh_shape_t* shape, coh_shape_t* minshape, const char* who) {
  int wx0 = floordiv(px0, 32) * 32, nw = (px1 - wx0) / 32 + 1, n_rows = py1 - py0 + 1;
  size_t nwords = (size_t)n_rows * nw;
  uint32_t *S = nullptr, *C = nullptr;
  CK(DMALLOC(&S, sizeof(uint32_t) * nwords)); CK(DMALLOC(&C, sizeof(uint32_t) * nwords));
  CK(cudaMemsetAsync(S, 0, sizeof(uint32_t) * nwords, ctx->stream));
  CK(cudaMemsetAsync(C, 0, sizeof(uint32_t) * nwords, ctx->stream));
  k_scan_rows<<<dim3(cdiv(n_rows, 64), cdiv(nw, SCAN_CHUNK_WORDS)), 64, 0, ctx->stream>>>(d_edges, n_edges, winding, py0, n_rows, wx0, nw, S, C, ctx->d_error); LAUNCHED();
  int rc = check_error_flag(ctx, who);
  if (!rc) rc = shape_from_bits(ctx, S, py0, n_rows, wx0, nw, shape);
  if (!rc) { k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(S, C, C, nwords, 1); LAUNCHED(); }  // minshape = shape - C
  if (!rc) rc = shape_from_bits(ctx, C, py0, n_rows, wx0, nw, minshape);
  DFREE(S); DFREE(C);
  return rc;
}
int coh_shapeminshape_of_edgelist(coh_ctx* ctx, const int32_t* edges, int32_t n_edges, int32_t winding,
                                  coh_shape_t* shape, coh_shape_t* minshape) {
  CK(cudaSetDevice(ctx->device));
  *shape = 0; *minshape = 0;
  if (n_edges <= 0) return 0;  // polygon.ml:584: NullShape, NullShape
  if (winding != COH_NONZERO && winding != COH_EVENODD) FAIL("bad winding rule");
  EdgeBox eb = edge_bounds(edges, n_edges);
  int px0, py0, px1, py1; shape_pixel_box(eb, px0, py0, px1, py1);
  EdgeRec* d_edges = nullptr;
  if (upload_edges(ctx, edges, n_edges, &d_edges)) return 1;
  int rc = shapes_from_device_edges(ctx, d_edges, n_edges, winding, px0, py0, px1, py1, shape, minshape, "coh_shapeminshape_of_edgelist");
  DFREE(d_edges);
  return rc;
}

// dense AA opacity bytes over the bit-frame of `shp`, then gathered in span order
static int polygon_opacity_dense(coh_ctx* ctx, const int32_t* edges, int n_edges, int winding, const DevShape* s,
                                 uint8_t** dense, int* wx0_out, int* nw_out) {
  int wx0 = floordiv(s->bx0, 32) * 32, nw = (s->bx1 - wx0) / 32 + 1;
  uint32_t* Q = nullptr;
  if (bits_from_shape(ctx, s, s->y0, s->n_rows, wx0, nw, &Q)) return 1;
  EdgeRec* d_edges = nullptr;
  if (upload_edges(ctx, edges, n_edges, &d_edges)) return 1;
  CK(DMALLOC(dense, (size_t)s->n_rows * nw * 32));
  CK(cudaMemsetAsync(*dense, 0, (size_t)s->n_rows * nw * 32, ctx->stream));
  dim3 g(cdiv(nw, 8), s->n_rows);
  k_aa_rows<<<g, 256, 0, ctx->stream>>>(d_edges, n_edges, winding, Q, s->y0, s->n_rows, wx0, nw, ctx->d_aa, *dense, ctx->d_error); LAUNCHED();
  int rc = check_error_flag(ctx, "coh_polygon_opacity");
  DFREE(Q); DFREE(d_edges);
  *wx0_out = wx0; *nw_out = nw;
  return rc;
}
int coh_polygon_opacity(coh_ctx* ctx, const int32_t* edges, int32_t n_edges, int32_t winding, coh_shape_t shp,
                        uint8_t* out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  if (!shp) return 0;
  DevShape* s = (DevShape*)shp;
  if (s->card > cap) FAIL("coh_polygon_opacity: buffer too small");
  uint8_t* dense = nullptr; int wx0, nw;
  if (n_edges <= 0) { memset(out, 0, (size_t)s->card); *n_out = s->card; return 0; }  // empty scaled shape: coverage 0
  if (polygon_opacity_dense(ctx, edges, n_edges, winding, s, &dense, &wx0, &nw)) return 1;
  std::vector<int> ptr; std::vector<int2> spans;
  if (download_shape(ctx, s, ptr, spans)) return 1;
  std::vector<uint8_t> h((size_t)s->n_rows * nw * 32);
  CK(cudaMemcpyAsync(h.data(), dense, h.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(dense);
  int64_t k = 0;
  for (int r = 0; r < s->n_rows; r++)
    for (int q = ptr[r]; q < ptr[r + 1]; q++)
      for (int i = 0; i < spans[q].y; i++) out[k++] = h[(size_t)r * nw * 32 + (spans[q].x + i - wx0)];
  *n_out = k;
  return 0;
}
int coh_polygon_sprite(coh_ctx* ctx, const coh_object* fill, const int32_t* edges, int32_t n_edges, int32_t winding,
                       coh_shape_t shp, uint32_t* out, int64_t cap, int64_t* n_out) {
  // polygon.ml:729-746: per span, colour = dissolve (fillsingle x_spanstart y) opacity.  The fill is
  // evaluated by the same device routine as the walker through a one-object render of `shp`'s spans;
  // here the opacity comes from the AA kernel and the (cheap, per-span) fill lookup runs on the host
  // side of the ABI only for this export entry point.
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  if (!shp) return 0;
  DevShape* s = (DevShape*)shp;
  if (s->card > cap) FAIL("coh_polygon_sprite: buffer too small");
  std::vector<uint8_t> op((size_t)s->card);
  int64_t n = 0;
  if (coh_polygon_opacity(ctx, edges, n_edges, winding, shp, op.data(), s->card, &n)) return 1;
  std::vector<int> ptr; std::vector<int2> spans;
  if (download_shape(ctx, s, ptr, spans)) return 1;
  FillRec f; f.kind = fill->fill_kind; f.c0 = fill->colour0; f.c1 = fill->colour1; f.flags = fill->fill_flags;
  for (int i = 0; i < 6; i++) f.p[i] = fill->fparam[i];
  int64_t k = 0;
  for (int r = 0; r < s->n_rows; r++)
    for (int q = ptr[r]; q < ptr[r + 1]; q++) {
      uint32_t c = fill_lookup(f, spans[q].x, s->y0 + r);
      for (int i = 0; i < spans[q].y; i++, k++) out[k] = px_dissolve(c, op[k]);
    }
  *n_out = k;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Convolve (convolve.mli:28-40)
// ---------------------------------------------------------------------------------------
static int conv_taps(coh_ctx* ctx, int kind, int r, int** d_taps, int* total) {
  *d_taps = nullptr; *total = 0;
  if (kind != COH_CONV_GAUSSIAN) return 0;
  std::vector<int> taps;
  for (int i = -r; i <= r; i++) {  // Convolve.mkgaussian r (convolve.ml:60-70)
    double xr = (double)i / (double)r, yr = 0. / (double)r;
    int v = (int)((double)(4 * r * r) * (exp(-(xr * xr + yr * yr)) / 2.) + 0.5);
    taps.push_back(v); *total += v;
  }
  CK(DMALLOC(d_taps, sizeof(int) * taps.size()));
  CK(cudaMemcpyAsync(*d_taps, taps.data(), sizeof(int) * taps.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
static int shape_pixel_offsets(coh_ctx* ctx, const DevShape* s, std::vector<long long>& off) {
  std::vector<int> ptr; std::vector<int2> spans;
  if (download_shape(ctx, s, ptr, spans)) return 1;
  off.assign(s->n_rows + 1, 0);
  for (int r = 0; r < s->n_rows; r++) { long long n = 0; for (int q = ptr[r]; q < ptr[r + 1]; q++) n += spans[q].y; off[r + 1] = off[r] + n; }
  return 0;
}
// Convolve.convolve_sprite kernel sprite (convolve.ml:239-258): the sprite is (shape, one RGBA8 per pixel in
// span order); the result lives on bloat r r (shape) and is returned the same way.
int coh_convolve_sprite(coh_ctx* ctx, int32_t kernel_kind, int32_t r, coh_shape_t shape, const uint32_t* rgba_in,
                        coh_shape_t* out_shape, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *out_shape = 0; *n_out = 0;
  if ((kernel_kind != COH_CONV_UNIT && kernel_kind != COH_CONV_GAUSSIAN) || r <= 0) FAIL("Convolve.mkunit / Convolve.mkxy: Invalid_argument");
  if (!shape) return 0;  // NullSprite -> NullSprite
  DevShape* s = (DevShape*)shape;
  coh_shape_t R = 0;
  if (coh_shape_bloat(ctx, shape, r, r, &R)) return 1;
  DevShape* rs = (DevShape*)R;
  if (rs->card > cap) { coh_shape_free(ctx, R); FAIL("coh_convolve_sprite: buffer too small"); }
  // canvas = bounding box grown by 2r (Sprite.flatten_sprite border, convolve.ml:247)
  const int x0 = s->bx0 - 2 * r, y0 = s->by0 - 2 * r, w = s->bx1 - s->bx0 + 1 + 4 * r, h = s->by1 - s->by0 + 1 + 4 * r;
  const size_t npx = (size_t)w * h;
  uint32_t *A = nullptr, *X = nullptr, *d_in = nullptr, *d_out = nullptr; long long* d_off = nullptr; int* d_taps = nullptr; int total = 0;
  std::vector<long long> off;
  if (shape_pixel_offsets(ctx, s, off)) return 1;
  CK(DMALLOC(&A, 4 * npx)); CK(DMALLOC(&X, 4 * npx));
  CK(cudaMemsetAsync(A, 0, 4 * npx, ctx->stream));
  CK(DMALLOC(&d_in, 4 * (size_t)std::max<long long>(s->card, 1))); CK(DMALLOC(&d_off, sizeof(long long) * off.size()));
  CK(cudaMemcpyAsync(d_in, rgba_in, 4 * (size_t)s->card, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, ctx->stream));
  k_scatter_spans<uint32_t><<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, d_off, s->n_rows, s->y0 - y0, x0, w, d_in, A); LAUNCHED();
  if (conv_taps(ctx, kernel_kind, r, &d_taps, &total)) return 1;
  dim3 gp(cdiv(w, 128), h);
  k_conv_pass<<<gp, 128, 0, ctx->stream>>>(A, X, w, h, r, kernel_kind, d_taps, total, 0); LAUNCHED();
  k_conv_pass<<<gp, 128, 0, ctx->stream>>>(X, A, w, h, r, kernel_kind, d_taps, total, 1); LAUNCHED();
  // pick the result up on R (Sprite.pickup)
  std::vector<long long> roff;
  if (shape_pixel_offsets(ctx, rs, roff)) return 1;
  long long* d_roff = nullptr;
  CK(DMALLOC(&d_roff, sizeof(long long) * roff.size())); CK(DMALLOC(&d_out, 4 * (size_t)std::max<long long>(rs->card, 1)));
  CK(cudaMemcpyAsync(d_roff, roff.data(), sizeof(long long) * roff.size(), cudaMemcpyHostToDevice, ctx->stream));
  // k_gather_spans indexes dense rows from the shape's first row: pass the canvas rows starting at R's first row
  k_gather_spans<uint32_t><<<cdiv(rs->n_rows, 128), 128, 0, ctx->stream>>>(rs->row_ptr, rs->spans, d_roff, rs->n_rows, x0, w, A + (size_t)(rs->y0 - y0) * w, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)rs->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(A); DFREE(X); DFREE(d_in); DFREE(d_out); DFREE(d_off); DFREE(d_roff); DFREE(d_taps);
  *out_shape = R; *n_out = rs->card;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Scenes and rendering
// ---------------------------------------------------------------------------------------
int coh_scene_free(coh_ctx* ctx, coh_scene_t h) {
  CK(cudaSetDevice(ctx->device));
  DevScene* s = (DevScene*)h;
  if (!s) return 0;
  DFREE(s->objs); DFREE(s->leaves); DFREE(s->leaf_box); DFREE(s->edges); DFREE(s->points); DFREE(s->stamps);
  DFREE(s->rowedge_ptr); DFREE(s->rowedge_idx); DFREE(s->brush_ranges); DFREE(s->conv_bits); DFREE(s->conv_px);
  for (auto& g : s->group_shape) free_shape(ctx, g.second.shape);
  delete s;
  return 0;
}

// brush.ml:60-92: alpha of the Gaussian stamp of white at `opacity`.
static void brush_stamp(double radius, double opacity, std::vector<uint8_t>& out, int& r_out) {
  int intopacity = (int)(opacity * 255.), intr = (int)ceil(radius);
  int size = 2 * intr + 1;
  r_out = intr;
  size_t base = out.size();
  out.resize(base + (size_t)size * size);
  uint32_t white = 0xFFFFFFFFu;
  uint32_t c1 = px_dissolve(white, intopacity);
  for (int y = 0; y < size; y++)
    for (int x = 0; x < size; x++) {
      double xp = (double)(x - intr), yp = (double)(y - intr), rr = radius / 2.;
      double v = 255. * exp(-((xp / rr) * (xp / rr) + (yp / rr) * (yp / rr)));
      int vi = (int)(v * 1.);
      out[base + (size_t)y * size + x] = (uint8_t)(px_dissolve(c1, vi) >> 24);
    }
}

int coh_scene_create(coh_ctx* ctx, const coh_object* objs, int32_t n_objs, int32_t n_background, const int32_t* edges,
                     int32_t n_edges, const int32_t* points, int32_t n_points, coh_scene_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (n_objs < 0 || n_background < 0 || n_background > n_objs) FAIL("scene: bad object counts");
  // The scene list and the (pages @ background) list are each wrapped in an implicit root group:
  // render_frame renders them separately over the same update and composites the two results
  // with `over` (render.ml:1357-1365), which is exactly what two sibling groups do in one walk.
  std::vector<ObjRec> recs;
  std::vector<int> leaves;
  std::vector<uint8_t> stamps;
  std::vector<int> open;  // indices (into recs) of open groups
  std::vector<int> edge_obj((size_t)std::max(n_edges, 1), -1);  // owning path object of every edge
  std::vector<int> point_obj((size_t)std::max(n_points, 1), -1);  // owning brush object of every point
  struct ConvItem { int rec, kind, r; };
  std::vector<ConvItem> conv_list;
  size_t conv_words = 0, conv_pixels = 0;
  long long total_rows = 0, total_brush_rows = 0;
  ObjRec root; memset(&root, 0, sizeof root);
  root.kind = K_GROUP; root.pretrans = -1; root.depth = 0; root.flags = OF_ROOT_SCENE;
  recs.push_back(root); open.push_back(0);
  std::vector<int> rec_of_abi((size_t)std::max(n_objs, 1), -1);
  std::vector<int> group_last;
  std::vector<int64_t> ids;
  // filters and their reading-scene groups (include/coherence_b200.h, COH_FILTER_*)
  std::vector<DevScene::FilterRec> filters;
  std::vector<int> filter_read_abi;            // per filter: abi index of its reading-scene group, or -1
  std::map<int, std::pair<int, int>> reading;  // abi index of a reading-scene GROUP_BEGIN -> leaf range
  std::vector<int> open_reading;               // per open GROUP_BEGIN: abi index if it is a reading-scene group, else -1
  int cur_reading = -1, n_scene_leaves = -1, n_front_leaves = -1;
  for (int i = 0; i < n_objs; i++) {
    if (i == n_objs - n_background) {
      if (open.size() != 1 || cur_reading >= 0) FAIL("scene: unterminated group");
      if (n_scene_leaves < 0) n_scene_leaves = (int)leaves.size();
      n_front_leaves = (int)leaves.size();
      root.flags = OF_ROOT_BACKGROUND;
      recs.push_back(root); open[0] = (int)recs.size() - 1;
    }
    const coh_object& c = objs[i];
    if (c.kind == COH_OBJ_GROUP_END) {
      if (open_reading.empty()) FAIL("scene: GROUP_END without GROUP_BEGIN");
      const int rd = open_reading.back(); open_reading.pop_back();
      if (rd >= 0) { reading[rd].second = (int)leaves.size(); cur_reading = -1; continue; }
      if (open.size() <= 1) FAIL("scene: GROUP_END without GROUP_BEGIN");
      group_last.resize(recs.size(), -1);
      group_last[open.back()] = (int)recs.size() - 1;
      open.pop_back();
      continue;
    }
    if (c.kind == COH_OBJ_GROUP_BEGIN && c.filter_kind == COH_FILTER_READING_SCENE) {
      // members become direct members of the root list of their own pass (render.ml:1091 renders the list)
      if (open.size() != 1 || cur_reading >= 0 || i >= n_objs - n_background) FAIL("scene: reading-scene groups must be top-level members of the scene list");
      if (n_scene_leaves < 0) n_scene_leaves = (int)leaves.size();
      open_reading.push_back(i); cur_reading = i;
      reading[i] = std::make_pair((int)leaves.size(), (int)leaves.size());
      continue;
    }
    if (n_scene_leaves >= 0 && cur_reading < 0 && i < n_objs - n_background) FAIL("scene: reading-scene groups must come after every ordinary scene object");
    if (c.kind == COH_OBJ_FILTER) {
      if (open.size() != 1 || cur_reading >= 0 || i >= n_objs - n_background) FAIL("scene: filter objects must be top-level members of the scene list");
      if (c.first < 0 || c.count < 0 || (int64_t)c.first + c.count > n_edges) FAIL("scene: edge range out of bounds");
      if (c.winding != COH_NONZERO && c.winding != COH_EVENODD) FAIL("scene: bad winding rule");
      if (c.filter_kind < COH_FILTER_HOLE || c.filter_kind > COH_FILTER_SCENE) FAIL("scene: bad filter kind");
      if (c.fill_kind != COH_FILL_PLAIN) FAIL("scene: filter geometry with a fancy fill is not supported yet");
      if (c.dx || c.dy) FAIL("scene: translated filter objects are not supported yet");
      DevScene::FilterRec f; memset(&f, 0, sizeof f);
      f.abi = i; f.pos = (int)leaves.size(); f.kind = c.filter_kind; f.first = c.first; f.count = c.count; f.winding = c.winding; f.colour = c.colour0;
      if (c.filter_kind == COH_FILTER_BLUR) {
        f.kernel_kind = c.filter_kernel & 255; f.r = c.filter_kernel >> 8;
        if ((f.kernel_kind != COH_CONV_UNIT && f.kernel_kind != COH_CONV_GAUSSIAN) || f.r <= 0 || f.r > 64) FAIL("Convolve.mkunit / mkxy: bad kernel");
      }
      if (c.count == 0) continue;  // NullShape geometry: the filter touches nothing
      EdgeBox eb = edge_bounds(edges + 4 * (size_t)c.first, c.count);
      shape_pixel_box(eb, f.bx0, f.by0, f.bx1, f.by1);
      filters.push_back(f); filter_read_abi.push_back(c.filter_kind == COH_FILTER_SCENE ? c.first2 : -1);
      continue;
    }
    ObjRec o; memset(&o, 0, sizeof o);
    o.pretrans = c.pretrans; o.dx = c.dx; o.dy = c.dy;
    if (c.pretrans < -1 || c.pretrans > 255) FAIL("scene: pretrans out of range");
    o.depth = (int)open.size();
    if (o.depth > MAX_DEPTH) FAIL("scene: groups nested too deeply (MAX_DEPTH)");
    for (int d = 0; d < o.depth; d++) o.anc[d] = open[d];
    o.fill.kind = c.fill_kind; o.fill.c0 = c.colour0; o.fill.c1 = c.colour1; o.fill.flags = c.fill_flags;
    for (int k = 0; k < 6; k++) o.fill.p[k] = c.fparam[k];
    switch (c.kind) {
      case COH_OBJ_GROUP_BEGIN:
        o.kind = K_GROUP;
        if (o.depth >= MAX_DEPTH) FAIL("scene: groups nested too deeply (MAX_DEPTH)");
        recs.push_back(o); open.push_back((int)recs.size() - 1); open_reading.push_back(-1);
        rec_of_abi[i] = (int)recs.size() - 1;
        ids.resize(recs.size(), -1); ids.back() = c.id;
        continue;
      case COH_OBJ_PATH: {
        if (c.first < 0 || c.count < 0 || (int64_t)c.first + c.count > n_edges) FAIL("scene: edge range out of bounds");
        if (c.winding != COH_NONZERO && c.winding != COH_EVENODD) FAIL("scene: bad winding rule");
        if (c.sprite_winding < 0 || c.sprite_winding > 2) FAIL("scene: bad sprite winding rule");
        o.kind = K_PATH; o.winding = c.winding; o.aa_winding = c.sprite_winding ? c.sprite_winding - 1 : c.winding;
        o.first = c.first; o.count = c.count;
        if (c.count == 0) continue;  // NullShape: nothing to draw
        if (c.convolve) {
          const int ck = c.convolve & 255, cr = c.convolve >> 8;
          if ((ck != COH_CONV_UNIT && ck != COH_CONV_GAUSSIAN) || cr <= 0 || cr > 64) FAIL("Convolve.mkunit / mkxy: bad kernel");  // convolve.ml:37-51 Invalid_argument
          if (c.fill_kind != COH_FILL_PLAIN) FAIL("scene: Convolved objects with fancy fills are not supported yet");
          o.kind = K_CONV;
          conv_list.push_back({(int)recs.size(), ck, cr});
        }
        EdgeBox eb = edge_bounds(edges + 4 * (size_t)c.first, c.count);
        shape_pixel_box(eb, o.bx0, o.by0, o.bx1, o.by1);
        if (o.kind == K_CONV) {  // the convolved object reaches r pixels further; its canvas another r (X-pass inputs)
          const int cr = c.convolve >> 8;
          o.cv_x0 = floordiv(o.bx0 - 2 * cr, 32) * 32; o.cv_y0 = o.by0 - 2 * cr;
          o.cv_nw = (o.bx1 + 2 * cr - o.cv_x0) / 32 + 1; o.cv_h = o.by1 + 2 * cr - o.cv_y0 + 1;
          o.bx0 -= cr; o.bx1 += cr; o.by0 -= cr; o.by1 += cr;
          o.cv_bits = (int)conv_words; conv_words += 2 * (size_t)o.cv_nw * o.cv_h;
          o.cv_px = (int)conv_pixels; conv_pixels += (size_t)o.cv_nw * 32 * o.cv_h;
          if (conv_words > 0x7FFFFFF0ull || conv_pixels > 0x7FFFFFF0ull) FAIL("scene: Convolved canvases too large");
        }
        // rows with a candidate edge list: extended band [32y-67, 32y+16] meets [ymin, ymax]
        o.ry0 = floordiv(eb.ymin - 16 + 31, 32); o.ry1 = floordiv(eb.ymax + 67, 32);
        if (total_rows + (o.ry1 - o.ry0 + 1) > 0x7FFFFFF0LL) FAIL("scene: too many object rows for the row-edge table");
        o.row_base = (int)total_rows; total_rows += o.ry1 - o.ry0 + 1;
        for (int k = 0; k < c.count; k++) {
          if (edge_obj[(size_t)c.first + k] != -1) FAIL("scene: objects may not share edges");
          edge_obj[(size_t)c.first + k] = (int)recs.size();
        }
        break;
      }
      case COH_OBJ_CPG: {
        if (c.first < 0 || c.count < 0 || (int64_t)c.first + c.count > n_edges) FAIL("scene: edge range out of bounds");
        if (c.first2 < 0 || c.count2 < 0 || (int64_t)c.first2 + c.count2 > n_edges) FAIL("scene: edge range out of bounds");
        if (c.first2 < c.first + c.count) FAIL("scene: CPG operand b's edges must follow operand a's");
        if ((c.winding != COH_NONZERO && c.winding != COH_EVENODD) || (c.winding2 != COH_NONZERO && c.winding2 != COH_EVENODD)) FAIL("scene: bad winding rule");
        if (c.cpg_op < COH_CPG_UNION || c.cpg_op > COH_CPG_EXCLUSIVEOR) FAIL("scene: bad CPG operator");
        if (c.convolve) FAIL("scene: Convolved CPG objects are not supported yet");
        o.kind = K_CPG; o.winding = o.aa_winding = c.winding;
        o.first = c.first; o.count = c.count; o.b_first = c.first2; o.b_count = c.count2; o.b_opw = c.cpg_op | (c.winding2 << 8);
        o.bx0 = o.by0 = INT32_MAX; o.bx1 = o.by1 = INT32_MIN;
        o.ry0 = o.b_ry0 = 0; o.ry1 = o.b_ry1 = -1;   // operands without edges have no rows
        for (int side = 0; side < 2; side++) {
          const int f = side ? c.first2 : c.first, n = side ? c.count2 : c.count;
          if (n == 0) continue;
          EdgeBox eb = edge_bounds(edges + 4 * (size_t)f, n);
          int x0, y0, x1, y1;
          shape_pixel_box(eb, x0, y0, x1, y1);
          o.bx0 = std::min(o.bx0, x0); o.by0 = std::min(o.by0, y0); o.bx1 = std::max(o.bx1, x1); o.by1 = std::max(o.by1, y1);
          const int r0 = floordiv(eb.ymin - 16 + 31, 32), r1 = floordiv(eb.ymax + 67, 32);
          if (total_rows + (r1 - r0 + 1) > 0x7FFFFFF0LL) FAIL("scene: too many object rows for the row-edge table");
          if (side) { o.b_ry0 = r0; o.b_ry1 = r1; o.b_row_base = (int)total_rows; } else { o.ry0 = r0; o.ry1 = r1; o.row_base = (int)total_rows; }
          total_rows += r1 - r0 + 1;
          for (int k = 0; k < n; k++) {
            if (edge_obj[(size_t)f + k] != -1) FAIL("scene: objects may not share edges");
            edge_obj[(size_t)f + k] = (int)recs.size();
          }
        }
        if (o.bx0 > o.bx1) continue;  // both operands null
        break;
      }
      case COH_OBJ_PRIMITIVE:
        o.kind = K_PRIM; o.fill.kind = 0;
        if (c.prim_null) continue;
        for (int k = 0; k < 4; k++) o.prim[k] = c.prim[k];
        if (c.prim[2] < c.prim[0] || c.prim[3] < c.prim[1]) FAIL("scene: primitive with negative extent");
        o.bx0 = c.prim[0]; o.by0 = c.prim[1]; o.bx1 = c.prim[2]; o.by1 = c.prim[3];
        break;
      case COH_OBJ_BRUSH: {
        if (c.first < 0 || c.count < 0 || (int64_t)c.first + c.count > n_points) FAIL("scene: point range out of bounds");
        if (!(c.brush_radius >= 0.) || !(c.brush_opacity >= 0. && c.brush_opacity <= 1.)) FAIL("scene: brush radius/opacity out of range");
        o.kind = K_BRUSH; o.first = c.first; o.count = c.count;
        if (c.count == 0) continue;
        o.stamp_off = (int)stamps.size();
        brush_stamp(c.brush_radius, c.brush_opacity, stamps, o.brush_r);
        int x0 = INT32_MAX, x1 = INT32_MIN, y0 = INT32_MAX, y1 = INT32_MIN;
        for (int k = 0; k < c.count; k++) {
          int px = points[2 * ((size_t)c.first + k)], py = points[2 * ((size_t)c.first + k) + 1];
          x0 = std::min(x0, px); x1 = std::max(x1, px); y0 = std::min(y0, py); y1 = std::max(y1, py);
        }
        o.bx0 = x0 - o.brush_r; o.bx1 = x1 + o.brush_r; o.by0 = y0 - o.brush_r; o.by1 = y1 + o.brush_r;
        o.ry0 = o.by0; o.ry1 = o.by1;   // object-frame rows (the alias offset is added to the box below)
        o.bc_x0 = floordiv(o.bx0, 32); o.bc_y0 = floordiv(o.by0, CELL_H);
        o.bc_nx = floordiv(o.bx1, 32) - o.bc_x0 + 1; o.bc_ny = floordiv(o.by1, CELL_H) - o.bc_y0 + 1;
        if (total_brush_rows + (long long)o.bc_nx * o.bc_ny > 0x7FFFFFF0LL) FAIL("scene: too many brush cells");
        o.bc_base = (int)total_brush_rows; total_brush_rows += (long long)o.bc_nx * o.bc_ny;
        for (int k = 0; k < c.count; k++) {
          if (point_obj[(size_t)c.first + k] != -1) FAIL("scene: objects may not share brush points");
          point_obj[(size_t)c.first + k] = (int)recs.size();
        }
        break;
      }
      default: FAIL("scene: unknown object kind");
    }
    o.bx0 += o.dx; o.bx1 += o.dx; o.by0 += o.dy; o.by1 += o.dy;
    if ((o.kind == K_PATH || o.kind == K_PRIM) && o.fill.kind == 0 && (o.fill.c0 >> 24) == 255u && o.pretrans < 0) {
      bool clear_path = true;
      for (int d = 0; d < o.depth; d++) clear_path = clear_path && recs[o.anc[d]].pretrans < 0;
      if (clear_path) o.flags |= OF_OCCLUDES;
    }
    recs.push_back(o);
    rec_of_abi[i] = (int)recs.size() - 1;
    ids.resize(recs.size(), -1); ids.back() = c.id;
    leaves.push_back((int)recs.size() - 1);
  }
  if (open.size() != 1 || cur_reading >= 0) FAIL("scene: unterminated group");
  if (n_scene_leaves < 0) n_scene_leaves = (int)leaves.size();
  if (n_front_leaves < 0) n_front_leaves = (int)leaves.size();
  for (size_t k = 0; k < filters.size(); k++) {
    if (filter_read_abi[k] < 0) continue;
    auto it = reading.find(filter_read_abi[k]);
    if (it == reading.end()) FAIL("scene: filter without its reading-scene group");
    filters[k].read0 = it->second.first; filters[k].read1 = it->second.second;
  }
  DevScene* s = new DevScene();
  s->filters = filters; s->n_scene_leaves = n_scene_leaves; s->n_front_leaves = n_front_leaves; s->h_leaves = leaves;
  s->n_objs = (int)recs.size(); s->n_leaves = (int)leaves.size(); s->n_edges = n_edges; s->n_points = n_points;
  s->h_objs = recs;
  group_last.resize(recs.size(), -1);
  ids.resize(recs.size(), -1);
  s->rec_of_abi = rec_of_abi; s->group_last = group_last; s->ids = ids;
  s->group_off.assign(recs.size(), make_int2(0, 0));
  for (const ObjRec& o : recs) {
    if (o.kind != K_GROUP && o.kind != K_PRIM && o.fill.kind != 0) s->has_fancy = true;
    if (o.kind == K_BRUSH || o.kind == K_CONV) s->extras = std::max(s->extras, 1);
    if (o.kind == K_CPG || !filters.empty()) s->extras = 2;  // the filter passes need the walker variant that can continue a frame
  }
  CK(DMALLOC(&s->objs, sizeof(ObjRec) * recs.size()));
  CK(cudaMemcpyAsync(s->objs, recs.data(), sizeof(ObjRec) * recs.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(DMALLOC(&s->leaves, sizeof(int) * std::max<size_t>(leaves.size(), 1)));
  if (!leaves.empty()) CK(cudaMemcpyAsync(s->leaves, leaves.data(), sizeof(int) * leaves.size(), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<int4> boxes(leaves.size());
  for (size_t i = 0; i < leaves.size(); i++) { const ObjRec& o = recs[leaves[i]]; boxes[i] = make_int4(o.bx0, o.by0, o.bx1, o.by1); }
  CK(DMALLOC(&s->leaf_box, sizeof(int4) * std::max<size_t>(leaves.size(), 1)));
  if (!leaves.empty()) CK(cudaMemcpyAsync(s->leaf_box, boxes.data(), sizeof(int4) * boxes.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (upload_edges(ctx, edges, n_edges, &s->edges)) return 1;
  CK(DMALLOC(&s->points, sizeof(int2) * std::max(n_points, 1)));
  if (n_points > 0) CK(cudaMemcpyAsync(s->points, points, sizeof(int2) * n_points, cudaMemcpyHostToDevice, ctx->stream));
  CK(DMALLOC(&s->stamps, std::max<size_t>(stamps.size(), 1)));
  if (!stamps.empty()) CK(cudaMemcpyAsync(s->stamps, stamps.data(), stamps.size(), cudaMemcpyHostToDevice, ctx->stream));
  // K1 edge binning: count -> scan -> fill
  {
    int* d_edge_obj = nullptr; int* d_counts = nullptr;
    size_t slots = (size_t)std::max<long long>(total_rows, 1);
    CK(DMALLOC(&d_edge_obj, sizeof(int) * edge_obj.size()));
    CK(cudaMemcpyAsync(d_edge_obj, edge_obj.data(), sizeof(int) * edge_obj.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(DMALLOC(&d_counts, sizeof(int) * slots));
    CK(DMALLOC(&s->rowedge_ptr, sizeof(int) * (slots + 1)));
    CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * slots, ctx->stream));
    if (n_edges > 0) { k_rowedges<false><<<cdiv(n_edges * 32, 256), 256, 0, ctx->stream>>>(s->edges, d_edge_obj, n_edges, s->objs, d_counts, nullptr, nullptr); LAUNCHED(); }
    if (exclusive_scan(ctx, d_counts, s->rowedge_ptr, (int)slots, nullptr)) return 1;
    // size of the lists: the same row range per edge as k_rowedges, summed on the host (no device round trip:
    // a device-to-host read here would queue behind an asynchronous framebuffer read-back of the previous frame)
    long long total = 0;
    for (int e = 0; e < n_edges; e++) {
      if (edge_obj[e] < 0) continue;
      const int ymin = std::min(edges[4 * (size_t)e + 1], edges[4 * (size_t)e + 3]), ymax = std::max(edges[4 * (size_t)e + 1], edges[4 * (size_t)e + 3]);
      total += floordiv(ymax + 67, 32) - floordiv(ymin - 16 + 31, 32) + 1;
    }
    if (total > 0x7FFFFFF0LL) FAIL("scene: row-edge table too large");
    CK(DMALLOC(&s->rowedge_idx, sizeof(int) * (size_t)std::max<long long>(total, 1)));
    CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * slots, ctx->stream));
    if (n_edges > 0) { k_rowedges<true><<<cdiv(n_edges * 32, 256), 256, 0, ctx->stream>>>(s->edges, d_edge_obj, n_edges, s->objs, d_counts, s->rowedge_ptr, s->rowedge_idx); LAUNCHED(); }
    CK(cudaStreamSynchronize(ctx->stream));
    DFREE(d_edge_obj); DFREE(d_counts);
  }
  // Convolved objects (render.ml:1023-1052): AA-rasterise the whole (twice bloated) box of the child,
  // X pass, Y pass; keep the shape / minshape bit-rows and the convolved canvas resident.
  if (!conv_list.empty()) {
    CK(DMALLOC(&s->conv_bits, sizeof(uint32_t) * conv_words));
    CK(DMALLOC(&s->conv_px, sizeof(uint32_t) * conv_pixels));
    for (const ConvItem& ci : conv_list) {
      const ObjRec& o = recs[ci.rec];
      const int nw = o.cv_nw, h = o.cv_h, w = nw * 32;
      const size_t nwords = (size_t)nw * h, npx = (size_t)w * h;
      uint32_t *S = nullptr, *C = nullptr, *T = nullptr, *Q = nullptr, *A = nullptr, *X = nullptr; uint8_t* op = nullptr; int* d_taps = nullptr;
      CK(DMALLOC(&S, 4 * nwords)); CK(DMALLOC(&C, 4 * nwords)); CK(DMALLOC(&T, 4 * nwords)); CK(DMALLOC(&Q, 4 * nwords));
      CK(DMALLOC(&A, 4 * npx)); CK(DMALLOC(&X, 4 * npx)); CK(DMALLOC(&op, npx));
      CK(cudaMemsetAsync(S, 0, 4 * nwords, ctx->stream)); CK(cudaMemsetAsync(C, 0, 4 * nwords, ctx->stream));
      CK(cudaMemsetAsync(op, 0, npx, ctx->stream));
      const EdgeRec* ed = s->edges + o.first;
      k_scan_rows<<<dim3(cdiv(h, 64), cdiv(nw, SCAN_CHUNK_WORDS)), 64, 0, ctx->stream>>>(ed, o.count, o.winding, o.cv_y0, h, o.cv_x0, nw, S, C, ctx->d_error); LAUNCHED();
      uint32_t* convS = s->conv_bits + o.cv_bits; uint32_t* convM = convS + nwords;
      dim3 g(cdiv(nw, 128), h);
      k_dilate<<<g, 128, 0, ctx->stream>>>(S, convS, h, nw, ci.r, ci.r); LAUNCHED();                  // shape = bloat r r (shape g)
      k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(S, C, C, nwords, 1); LAUNCHED();  // C := minshape g
      k_fill_words<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(Q, nwords, 0xFFFFFFFFu); LAUNCHED();
      k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(Q, C, T, nwords, 1); LAUNCHED();  // T := frame - minshape
      k_dilate<<<g, 128, 0, ctx->stream>>>(T, S, h, nw, ci.r, ci.r); LAUNCHED();                       // S := bloat (frame - minshape)
      k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(C, S, convM, nwords, 1); LAUNCHED();  // minshape = erode r r (minshape g)
      k_aa_rows<<<dim3(cdiv(nw, 8), h), 256, 0, ctx->stream>>>(ed, o.count, o.aa_winding, Q, o.cv_y0, h, o.cv_x0, nw, ctx->d_aa, op, ctx->d_error); LAUNCHED();
      k_raster_plain<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(op, A, npx, o.fill.c0); LAUNCHED();
      std::vector<int> taps; int total = 0;
      if (ci.kind == COH_CONV_GAUSSIAN) {  // Convolve.mkgaussian r (convolve.ml:60-70)
        for (int i = -ci.r; i <= ci.r; i++) {
          double xr = (double)i / (double)ci.r, yr = 0. / (double)ci.r;
          double gg = exp(-(xr * xr + yr * yr)) / 2.;
          int v = (int)((double)(4 * ci.r * ci.r) * gg + 0.5);
          taps.push_back(v); total += v;
        }
        CK(DMALLOC(&d_taps, sizeof(int) * taps.size()));
        CK(cudaMemcpyAsync(d_taps, taps.data(), sizeof(int) * taps.size(), cudaMemcpyHostToDevice, ctx->stream));
      }
      dim3 gp(cdiv(w, 128), h);
      k_conv_pass<<<gp, 128, 0, ctx->stream>>>(A, X, w, h, ci.r, ci.kind, d_taps, total, 0); LAUNCHED();
      k_conv_pass<<<gp, 128, 0, ctx->stream>>>(X, s->conv_px + o.cv_px, w, h, ci.r, ci.kind, d_taps, total, 1); LAUNCHED();
      if (check_error_flag(ctx, "coh_scene_create (Convolved object)")) return 1;
      DFREE(S); DFREE(C); DFREE(T); DFREE(Q); DFREE(A); DFREE(X); DFREE(op); DFREE(d_taps);
    }
  }
  if (total_brush_rows > 0) {
    int* d_point_obj = nullptr;
    CK(DMALLOC(&d_point_obj, sizeof(int) * point_obj.size()));
    CK(cudaMemcpyAsync(d_point_obj, point_obj.data(), sizeof(int) * point_obj.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(DMALLOC(&s->brush_ranges, sizeof(int2) * (size_t)total_brush_rows));
    std::vector<int2> init((size_t)total_brush_rows, make_int2(INT32_MAX, -1));
    CK(cudaMemcpyAsync(s->brush_ranges, init.data(), sizeof(int2) * init.size(), cudaMemcpyHostToDevice, ctx->stream));
    k_brush_cells<<<cdiv(n_points, 256), 256, 0, ctx->stream>>>(s->points, d_point_obj, n_points, s->objs, s->brush_ranges); LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));
    DFREE(d_point_obj);
  }
  *out = (coh_scene_t)s;
  return 0;
}

int coh_fb_attach(coh_ctx* ctx, void* device_rgba8) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fr.W) FAIL("coh_fb_attach: call coh_fb_configure first");
  if (drain_timing(ctx)) return 1;
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_fb) DFREE(ctx->fb);
  if (!device_rgba8) {  // detach: back to a framebuffer owned by the context
    ctx->fb = nullptr; ctx->own_fb = true;
    CK(DMALLOC(&ctx->fb, sizeof(uint32_t) * (size_t)ctx->fr.W * ctx->fr.H));
    CK(cudaMemsetAsync(ctx->fb, 0, sizeof(uint32_t) * (size_t)ctx->fr.W * ctx->fr.H, ctx->stream));
    return 0;
  }
  ctx->fb = (uint32_t*)device_rgba8; ctx->own_fb = false;
  return 0;
}
int coh_fb_set_peers(coh_ctx* ctx, int32_t n_peers, void* const* peer_fbs) {
  if (n_peers < 0 || n_peers > COH_MAX_PEERS) FAIL("coh_fb_set_peers: at most 7 peers (one 8-GPU box)");
  ctx->n_peers = n_peers;
  for (int k = 0; k < n_peers; k++) ctx->peer_fb[k] = (uint32_t*)peer_fbs[k];
  return 0;
}
int coh_fb_configure(coh_ctx* ctx, int32_t width, int32_t height, int32_t band_y0, int32_t band_y1) {
  CK(cudaSetDevice(ctx->device));
  if (width <= 0 || height <= 0) FAIL("coh_fb_configure: bad size");
  if (band_y0 < 0 || band_y1 > height || band_y0 > band_y1) FAIL("coh_fb_configure: bad band");
  if (width != ctx->fr.W || height != ctx->fr.H) {
    if (ctx->own_fb) DFREE(ctx->fb);
    DFREE(ctx->u_out); DFREE(ctx->u_init); ctx->fb = nullptr; ctx->u_out = nullptr; ctx->u_init = nullptr; ctx->own_fb = true;
    CK(DMALLOC(&ctx->fb, sizeof(uint32_t) * (size_t)width * height));
    CK(cudaMemsetAsync(ctx->fb, 0, sizeof(uint32_t) * (size_t)width * height, ctx->stream));
    CK(DMALLOC(&ctx->u_out, sizeof(uint32_t) * (size_t)cdiv(width, 32) * height));
  }
  ctx->fr.W = width; ctx->fr.H = height; ctx->fr.band_y0 = band_y0; ctx->fr.band_y1 = band_y1;
  ctx->fr.tiles_x = cdiv(width, 32); ctx->fr.cells_y = cdiv(height, CELL_H);
  ctx->fr.ctx0 = 0; ctx->fr.cntx = ctx->fr.tiles_x;
  ctx->have_u = false;
  return 0;
}

// One walk over the leaves [l0, l1) of a scene: binning + k_walk.
struct PassArgs {
  int l0, l1;                 // leaf range (list order)
  int ux, uy, uw, uh;         // update box (used when u_init is null)
  const uint32_t* u_init;     // update set as a bit-frame, or null
  uint32_t* u_out;            // receives `u` after the scene list, or null (may alias u_init)
  uint32_t* fb;               // target canvas
  bool write_clear, resume;
};
static int render_pass(coh_ctx* ctx, DevScene* s, const PassArgs& A) {
  Frame fr = ctx->fr;
  const int ux = A.ux, uy = A.uy, uw = A.uw, uh = A.uh;
  const bool write_clear = A.write_clear;
  const int n_leaves = A.l1 - A.l0;
  const int4* leaf_box = s->leaf_box + A.l0;
  const int* leaves = s->leaves + A.l0;
  if (fr.band_y1 <= fr.band_y0 || uw <= 0 || uh <= 0) return 0;
  // only the cell rows the update box reaches (a dirty region is usually a small part of the frame)
  const int ry0 = std::max(fr.band_y0, uy), ry1 = std::min(fr.band_y1, uy + uh);
  if (ry1 <= ry0) return 0;
  // ... and only the tile columns it reaches
  fr.ctx0 = std::max(0, ux >> 5);
  const int ctx1 = std::min(fr.tiles_x - 1, (int)(((long long)ux + uw - 1) >> 5));
  if (ctx1 < fr.ctx0) return 0;
  fr.cntx = ctx1 - fr.ctx0 + 1;
  const bool whole = A.l0 == 0 && A.l1 == s->n_leaves && ry0 == fr.band_y0 && ry1 == fr.band_y1 && fr.cntx == fr.tiles_x;
  int cell_row0 = ry0 / CELL_H, cell_row1 = (ry1 - 1) / CELL_H;
  if (A.u_out && A.u_out != A.u_init && !(ry0 == fr.band_y0 && ry1 == fr.band_y1 && fr.cntx == fr.tiles_x))
    // rows and columns the walk does not visit have nothing uncovered
    CK(cudaMemsetAsync(A.u_out + (size_t)fr.band_y0 * fr.tiles_x, 0, 4 * (size_t)(fr.band_y1 - fr.band_y0) * fr.tiles_x, ctx->stream));
  int n_cells = (cell_row1 - cell_row0 + 1) * fr.cntx;
  if (n_cells > ctx->n_cells_cap) {
    DFREE(ctx->cell_order); DFREE(ctx->cell_head);
    CK(DMALLOC(&ctx->cell_head, sizeof(int2) * n_cells));
    DFREE(ctx->cell_rng);
    CK(DMALLOC(&ctx->cell_rng, sizeof(int2) * n_cells));
    CK(DMALLOC(&ctx->cell_order, sizeof(int) * (size_t)n_cells * BIN_CLASSES));  // one-pass binning keeps one segment per length class
    ctx->n_cells_cap = n_cells;
  }
  if (!ctx->queue) {
    CK(DMALLOC(&ctx->order_hist, sizeof(int) * (2 * ORDER_BINS + 1)));  // histogram, cursors, work-queue head
    ctx->queue = ctx->order_hist + 2 * ORDER_BINS;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->n_sms = prop.multiProcessorCount;
  }
  const bool ordered = !s->has_fancy;  // with fancy fills the queue must stay row-major (carry look-back)
  CK(cudaMemsetAsync(ctx->order_hist, 0, sizeof(int) * (2 * ORDER_BINS + 1), ctx->stream));
  if (ctx->timing) { if (drain_timing(ctx)) return 1; CK(cudaEventRecord(ctx->ev[0], ctx->stream)); }
  // K1: count, scan, fill.  Small scenes: warp per cell scanning all leaves (lists come out sorted,
  // no atomics).  Large scenes: warp per leaf over the cells it covers + per-cell sort.
  const bool big = n_leaves > 1024;
  // capacity of the item pool: the exact total is a pure function of the object boxes and the
  // frame geometry, so it is computed on the host (once per scene and geometry) — no device
  // round trip inside a frame.
  size_t total = 0;
  if (!whole || s->items_for_W != fr.W || s->items_for_H != fr.H || s->items_for_y0 != fr.band_y0 || s->items_for_y1 != fr.band_y1) {
    size_t tot = 0;
    for (int li = A.l0; li < A.l1; li++) {
      const ObjRec& o = s->h_objs[s->h_leaves[li]];
      int cx0 = std::max(o.bx0 >> 5, fr.ctx0), cx1 = std::min(o.bx1 >> 5, ctx1);
      int cy0 = std::max(floordiv(o.by0, CELL_H), cell_row0), cy1 = std::min(floordiv(o.by1, CELL_H), cell_row1);
      if (cx1 >= cx0 && cy1 >= cy0) tot += (size_t)(cx1 - cx0 + 1) * (cy1 - cy0 + 1);
    }
    total = tot;
    if (whole) { s->coarse_total_valid = false; s->items_total = tot; s->items_for_W = fr.W; s->items_for_H = fr.H; s->items_for_y0 = fr.band_y0; s->items_for_y1 = fr.band_y1; }
  } else total = s->items_total;
  const size_t need = total;
  if (need > ctx->cell_items_cap) {
    DFREE(ctx->cell_items); DFREE(ctx->item_cell);
    size_t cap = need + need / 2 + 1024;
    CK(DMALLOC(&ctx->cell_items, sizeof(int) * cap));
    CK(DMALLOC(&ctx->item_cell, sizeof(int) * cap));
    ctx->cell_items_cap = cap;
  }
  if (!big) {
    // one pass: hit masks in registers, lists carved from one cursor, length classes instead of a sort
    const int bin_blocks = cdiv(n_cells * 32, 256);
    BinPrefill pf; memset(&pf, 0, sizeof pf);
    // (not with peer framebuffers: the mirrored stores of background cells are better spread over the walker's
    // warps — measured at 2 / 4 / 8 GPUs)
    if (ordered && !A.u_init && !A.resume && !(A.fb == ctx->fb && ctx->n_peers > 0)) {
      pf.fb = A.fb; pf.u_out = A.u_out; pf.ux0 = ux; pf.uy0 = uy; pf.ux1 = ux + uw - 1; pf.uy1 = uy + uh - 1;
      pf.n_peers = 0;
    }
    k_bin1<<<bin_blocks, 256, 0, ctx->stream>>>(leaf_box, leaves, n_leaves, fr, cell_row0, n_cells, ctx->cell_rng, ctx->cell_items, ctx->order_hist,
                                               ordered ? ctx->cell_order : nullptr, s->objs, ctx->cell_head, ctx->item_cell, pf); LAUNCHED();
  } else {
    // two levels: leaves into coarse cells (object-parallel, sorted per coarse list), then every fine cell from its coarse list
    const int ctx_x = cdiv(fr.tiles_x, COARSE), crow0 = cell_row0 >> COARSE_SHIFT, crow1 = cell_row1 >> COARSE_SHIFT;
    const int n_coarse = ctx_x * (crow1 - crow0 + 1);
    size_t ctot = 0;
    if (whole && s->coarse_total_valid) ctot = s->coarse_total;   // a pure function of the boxes and the frame geometry, like items_total
    else {
      for (int li = A.l0; li < A.l1; li++) {
        const ObjRec& o = s->h_objs[s->h_leaves[li]];
        int cx0 = std::max(floordiv(o.bx0, 32 * COARSE), 0), cx1 = std::min(floordiv(o.bx1, 32 * COARSE), ctx_x - 1);
        int cy0 = std::max(floordiv(o.by0, CELL_H * COARSE), crow0), cy1 = std::min(floordiv(o.by1, CELL_H * COARSE), crow1);
        if (cx1 >= cx0 && cy1 >= cy0) ctot += (size_t)(cx1 - cx0 + 1) * (cy1 - cy0 + 1);
      }
      if (whole) { s->coarse_total = ctot; s->coarse_total_valid = true; }
    }
    if (2 * ctot + 1 > ctx->coarse_cap || (size_t)n_coarse + 1 > ctx->coarse_cells_cap) {
      DFREE(ctx->coarse_items); DFREE(ctx->coarse_counts); DFREE(ctx->coarse_off);
      ctx->coarse_cap = 2 * ctot + ctot / 2 + 1024; ctx->coarse_cells_cap = (size_t)n_coarse + 1;
      CK(DMALLOC(&ctx->coarse_items, sizeof(int) * ctx->coarse_cap));
      CK(DMALLOC(&ctx->coarse_counts, sizeof(int) * ctx->coarse_cells_cap));
      CK(DMALLOC(&ctx->coarse_off, sizeof(int) * (ctx->coarse_cells_cap + 1)));
    }
    const int obj_blocks = cdiv(std::max(n_leaves, 1) * 32, 256);
    CK(cudaMemsetAsync(ctx->coarse_counts, 0, sizeof(int) * n_coarse, ctx->stream));
    k_bin_obj<false><<<obj_blocks, 256, 0, ctx->stream>>>(leaf_box, n_leaves, ctx_x, crow0, crow1, ctx->coarse_counts, nullptr, nullptr); LAUNCHED();
    if (exclusive_scan(ctx, ctx->coarse_counts, ctx->coarse_off, n_coarse, nullptr)) return 1;
    CK(cudaMemsetAsync(ctx->coarse_counts, 0, sizeof(int) * n_coarse, ctx->stream));
    k_bin_obj<true><<<obj_blocks, 256, 0, ctx->stream>>>(leaf_box, n_leaves, ctx_x, crow0, crow1, ctx->coarse_counts, ctx->coarse_off, ctx->coarse_items); LAUNCHED();
    k_bin_sort<<<cdiv(n_coarse * 32, 128), 128, 0, ctx->stream>>>(ctx->coarse_off, ctx->coarse_items, ctx->coarse_items + ctot, n_coarse); LAUNCHED();
    k_bin2<<<cdiv(n_cells * 32, 256), 256, 0, ctx->stream>>>(leaf_box, leaves, ctx->coarse_off, ctx->coarse_items, ctx_x, crow0, fr, cell_row0, n_cells, ctx->cell_rng,
                                                          ctx->cell_items, ctx->order_hist, ordered ? ctx->cell_order : nullptr); LAUNCHED();
  }
  WalkParams P;
  P.objs = s->objs; P.edges = s->edges; P.points = s->points; P.stamps = s->stamps;
  P.rowedge_ptr = s->rowedge_ptr; P.rowedge_idx = s->rowedge_idx; P.brush_ranges = s->brush_ranges;
  P.conv_bits = s->conv_bits; P.conv_px = s->conv_px;
  P.cell_rng = ctx->cell_rng;
  P.cls_cells = ordered ? ctx->cell_order : nullptr; P.cls_cnt = ordered ? ctx->order_hist + 1 : nullptr;
  P.cell_items = ctx->cell_items; P.cell_head = big ? nullptr : ctx->cell_head; P.aa = ctx->d_aa; P.fr = fr; P.cell_row0 = cell_row0;
  P.ux0 = ux; P.uy0 = uy; P.ux1 = ux + uw - 1; P.uy1 = uy + uh - 1;
  P.u_init = A.u_init; P.u_out = A.u_out; P.fb = A.fb; P.error_flag = ctx->d_error;
  P.write_clear = write_clear ? 1 : 0; P.resume = A.resume ? 1 : 0;
  P.n_peers = (A.fb == ctx->fb) ? ctx->n_peers : 0;   // only the frame itself is mirrored, not filter canvases
  for (int k = 0; k < COH_MAX_PEERS; k++) P.peer_fb[k] = k < P.n_peers ? ctx->peer_fb[k] : nullptr;
  // persistent grid: exactly one resident wave.  Work items are 4 rows high, or 16 for very large scenes.
  int walk_h = big ? 16 : 4;
  // Few cells (a band of an 8-GPU split, a small dirty region): the launch is bounded by its longest work item,
  // not by throughput — one-row items shorten that path (measured on the lion at 8 GPUs: 0.164 -> 0.141 ms;
  // at 1 to 4 GPUs four-row items are as fast or faster).
  if (!big && (long long)n_cells * 4 < 3LL * ctx->n_sms * WALK_MIN_CTAS * WALK_WARPS) walk_h = 1;
  if (const char* e = getenv("COH_WALK_H")) { int v = atoi(e); if (v == 1 || v == 4 || v == 16) walk_h = v; }  // tests force every variant
  const int grid = std::min(ctx->n_sms * WALK_MIN_CTAS, cdiv(n_cells * (CELL_H / walk_h), WALK_WARPS));
#define LAUNCH_WALK_E(CARRYV, EX)                                                                                  \
  do {                                                                                                             \
    if (walk_h == 4) k_walk<CARRYV, EX, 4><<<grid, WALK_WARPS * 32, 0, ctx->stream>>>(P);                         \
    else if (walk_h == 1) k_walk<CARRYV, EX, 1><<<grid, WALK_WARPS * 32, 0, ctx->stream>>>(P);                    \
    else k_walk<CARRYV, EX, 16><<<grid, WALK_WARPS * 32, 0, ctx->stream>>>(P);                                    \
    LAUNCHED();                                                                                                    \
  } while (0)
#define LAUNCH_WALK(CARRYV)                                                                                        \
  do {                                                                                                             \
    if (s->extras == 0) LAUNCH_WALK_E(CARRYV, 0); else if (s->extras == 1) LAUNCH_WALK_E(CARRYV, 1); else LAUNCH_WALK_E(CARRYV, 2); \
  } while (0)
  P.queue = ctx->queue; P.n_cells = n_cells;
  P.pre_sc = nullptr; P.pre_op = nullptr; P.item_cell = nullptr;
  if (ctx->timing) CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  // Plain-filled paths and primitives only, a list pool of moderate size: three-phase frame (kernels.cuh)
  // (small launches — a band of an 8-GPU split, a dirty region — stay fused: four dependent launches cost more
  // than the parallelism gains there; measured on 1/8 bands of the lion: 0.051 vs 0.059 ms)
  const char* force = getenv("COH_FUSED");   // tests force either path: "1" fused, "0" three-phase
  const bool pre = s->extras == 0 && !A.resume && !big && total > 0 && total * CELL_H <= (size_t)(1 << 23) &&
                   (force ? force[0] == '0' : walk_h != 1);
  if (pre) {
    const size_t n_pairs = total * CELL_H;
    if (n_pairs > ctx->pre_cap) {
      DFREE(ctx->pre_sc); DFREE(ctx->pre_list); DFREE(ctx->pre_op);
      const size_t cap = n_pairs + n_pairs / 4 + 1024;
      CK(DMALLOC(&ctx->pre_sc, sizeof(uint2) * cap));
      CK(DMALLOC(&ctx->pre_list, sizeof(int4) * cap)); CK(DMALLOC(&ctx->pre_op, 32 * cap));
      if (!ctx->pre_n) CK(DMALLOC(&ctx->pre_n, sizeof(int)));
      ctx->pre_cap = cap;
    }
    P.item_cell = ctx->item_cell;
    CK(cudaMemsetAsync(ctx->pre_n, 0, sizeof(int), ctx->stream));
    k_pre_scan<<<cdiv((int)n_pairs, 128), 128, 0, ctx->stream>>>(P, (int)n_pairs, ctx->pre_sc); LAUNCHED();
    k_pre_vis<<<cdiv(n_cells * CELL_H, 128), 128, 0, ctx->stream>>>(P, ctx->pre_sc, ctx->pre_list, ctx->pre_n); LAUNCHED();
    k_pre_aa<<<ctx->n_sms * 4, 256, 0, ctx->stream>>>(P, ctx->pre_list, ctx->pre_n, ctx->pre_op); LAUNCHED();
    P.pre_sc = ctx->pre_sc; P.pre_op = ctx->pre_op;
    const int pgrid = std::min(ctx->n_sms * WALK_MIN_CTAS, cdiv(n_cells * (CELL_H / 4), WALK_WARPS));
    if (s->has_fancy) {  // fancy fills: the compositing walk keeps the cross-tile carry (row-major queue order)
      size_t slots = (size_t)fr.tiles_x * (fr.band_y1 - fr.band_y0);
      if (slots > ctx->carry_slots) {
        DFREE(ctx->carry_done); DFREE(ctx->carry_cnt); DFREE(ctx->carry_ent);
        CK(DMALLOC(&ctx->carry_done, sizeof(int) * slots));
        CK(DMALLOC(&ctx->carry_cnt, sizeof(int) * slots));
        CK(DMALLOC(&ctx->carry_ent, sizeof(int2) * slots * CARRY_CAP));
        CK(cudaMemsetAsync(ctx->carry_done, 0, sizeof(int) * slots, ctx->stream));
        ctx->carry_slots = slots;
      }
      P.carry_done = ctx->carry_done; P.carry_cnt = ctx->carry_cnt; P.carry_ent = ctx->carry_ent;
      P.epoch = ++ctx->epoch;
      k_walk<true, 0, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P); LAUNCHED();
    } else {
      k_walk<false, 0, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P); LAUNCHED();
    }
    if (ctx->timing) { CK(cudaEventRecord(ctx->ev[2], ctx->stream)); ctx->ev_pending = true; }
    return 0;
  }
  P.carry_done = nullptr; P.carry_cnt = nullptr; P.carry_ent = nullptr; P.epoch = 0;
  if (s->has_fancy) {
    size_t slots = (size_t)fr.tiles_x * (fr.band_y1 - fr.band_y0);
    if (slots > ctx->carry_slots) {
      DFREE(ctx->carry_done); DFREE(ctx->carry_cnt); DFREE(ctx->carry_ent);
      CK(DMALLOC(&ctx->carry_done, sizeof(int) * slots));
      CK(DMALLOC(&ctx->carry_cnt, sizeof(int) * slots));
      CK(DMALLOC(&ctx->carry_ent, sizeof(int2) * slots * CARRY_CAP));
      CK(cudaMemsetAsync(ctx->carry_done, 0, sizeof(int) * slots, ctx->stream));
      ctx->carry_slots = slots;
    }
    P.carry_done = ctx->carry_done; P.carry_cnt = ctx->carry_cnt; P.carry_ent = ctx->carry_ent;
    P.epoch = ++ctx->epoch;
    LAUNCH_WALK(true);
  } else {
    LAUNCH_WALK(false);
  }
  if (ctx->timing) { CK(cudaEventRecord(ctx->ev[2], ctx->stream)); ctx->ev_pending = true; }
  return 0;
}

// ---------------------------------------------------------------------------------------
// Frames with filter objects (render.ml:1080-1131, 1248-1265; filters.ml).  A filter splits the
// scene list: the members in front of it are walked as usual; the filter itself renders its
// reading scene (X) and the members below it (Z) into canvases of their own — each a recursive
// render of the rest of the list, as in the reference — filters X, and blends the two by the
// antialiased matte of its geometry into the accumulator; its whole shape then leaves `u`
// (the "extra finish", render.ml:1120-1121, 1308) and the walk continues below it, the
// accumulator carrying on from the framebuffer (WalkParams::resume).
// ---------------------------------------------------------------------------------------
struct PixBox { int x0, y0, x1, y1; };   // inclusive pixel box; empty when x1 < x0 or y1 < y0
static int render_suffix(coh_ctx* ctx, DevScene* s, int l0, int f0, uint32_t* U, uint32_t* target, bool fresh, PixBox box, bool target_zeroed = false);

// `box` bounds the set bits of U: all work is confined to its rows (bit-frames are small and handled
// whole; the RGBA8 canvases are only touched in the rows the filter reads or writes).
static int apply_filter(coh_ctx* ctx, DevScene* s, int fi, uint32_t* U, uint32_t* target, PixBox box) {
  const DevScene::FilterRec& F = s->filters[fi];
  const Frame& fr = ctx->fr;
  const int W = fr.W, H = fr.H, nw = fr.tiles_x;
  const size_t nwords = (size_t)nw * H;
  // rows / columns of shptorender = shape(geometry) ∩ u
  const int y0 = std::max(std::max(F.by0, 0), box.y0), y1 = std::min(std::min(F.by1, H - 1), box.y1);
  const int x0 = std::max(std::max(F.bx0, 0), box.x0), x1 = std::min(std::min(F.bx1, W - 1), box.x1);
  if (y0 > y1 || x0 > x1) return 0;  // the geometry cannot meet u: nothing to render, nothing leaves u
  const int h = y1 - y0 + 1;
  const int m = F.kind == COH_FILTER_BLUR ? 2 * F.r + 1 : 0;             // reach of the reading shape
  const int ry0 = std::max(0, y0 - m), ry1 = std::min(H - 1, y1 + m), rh = ry1 - ry0 + 1;
  const PixBox tbox{x0, y0, x1, y1}, rbox{std::max(0, x0 - m), ry0, std::min(W - 1, x1 + m), ry1};
  const unsigned wblocks = (unsigned)((nwords + 255) / 256);
  const size_t po = (size_t)ry0 * W, pn = (size_t)rh * W;                 // canvas rows [ry0, ry1]
  uint32_t *SG = nullptr, *CG = nullptr, *T = nullptr, *R = nullptr, *X = nullptr, *Z = nullptr, *tmp = nullptr;
  uint8_t *op = nullptr, *alpha = nullptr; int* d_taps = nullptr;
  CK(DMALLOC(&SG, 4 * nwords)); CK(DMALLOC(&CG, 4 * nwords)); CK(DMALLOC(&T, 4 * nwords)); CK(DMALLOC(&R, 4 * nwords));
  CK(cudaMemsetAsync(SG, 0, 4 * nwords, ctx->stream)); CK(cudaMemsetAsync(CG, 0, 4 * nwords, ctx->stream));
  const EdgeRec* ed = s->edges + F.first;
  // shape of the geometry (render.ml:472-474); CG receives the coverage (the minshape is needed for the matte)
  k_scan_rows<<<dim3(cdiv(h, 64), cdiv(nw, SCAN_CHUNK_WORDS)), 64, 0, ctx->stream>>>(ed, F.count, F.winding, y0, h, 0, nw, SG + (size_t)y0 * nw, CG + (size_t)y0 * nw, ctx->d_error); LAUNCHED();
  k_bitop<<<wblocks, 256, 0, ctx->stream>>>(SG, U, T, nwords, 2); LAUNCHED();     // shptorender = r &&& u (render.ml:1281)
  // reading scene -> X -> filter function -> Y (in place)
  uint32_t* Y = nullptr;
  if (F.kind != COH_FILTER_HOLE) {
    CK(DMALLOC(&X, 4 * (size_t)W * H));
    CK(cudaMemsetAsync(X + po, 0, 4 * pn, ctx->stream));
    if (F.kind == COH_FILTER_BLUR) {  // filters.ml:247-250: read in bloat (2r+1) (2r+1) shp
      CK(cudaMemsetAsync(R, 0, 4 * nwords, ctx->stream));
      k_dilate<<<dim3(cdiv(nw, 128), rh), 128, 0, ctx->stream>>>(T + (size_t)ry0 * nw, R + (size_t)ry0 * nw, rh, nw, m, m); LAUNCHED();  // T is empty outside [y0, y1]
    } else CK(cudaMemcpyAsync(R, T, 4 * nwords, cudaMemcpyDeviceToDevice, ctx->stream));
    if (F.kind == COH_FILTER_SCENE) {
      PassArgs A{F.read0, F.read1, rbox.x0, rbox.y0, rbox.x1 - rbox.x0 + 1, rbox.y1 - rbox.y0 + 1, R, nullptr, X, true, false};
      if (render_pass(ctx, s, A)) return 1;
    } else if (render_suffix(ctx, s, F.pos, fi + 1, R, X, true, rbox, true)) return 1;
    Y = X;
    if (F.kind == COH_FILTER_MONOCHROME) {
      k_monochrome<<<(unsigned)((pn + 255) / 256), 256, 0, ctx->stream>>>(X + po, X + po, pn); LAUNCHED();
    } else if (F.kind == COH_FILTER_BLUR) {
      // Convolve.convolve_sprite_in_shape (convolve.ml:265-296) on the canvas rows [ry0, ry1]: pixels the
      // reading scene did not render are clear, exactly like the reference's canvas outside the sprite
      std::vector<int> taps; int total = 0;
      if (F.kernel_kind == COH_CONV_GAUSSIAN) {  // Convolve.mkgaussian r (convolve.ml:60-70)
        for (int i = -F.r; i <= F.r; i++) {
          double xr = (double)i / (double)F.r, yr = 0. / (double)F.r;
          double gg = exp(-(xr * xr + yr * yr)) / 2.;
          int v = (int)((double)(4 * F.r * F.r) * gg + 0.5);
          taps.push_back(v); total += v;
        }
        CK(DMALLOC(&d_taps, sizeof(int) * taps.size()));
        CK(cudaMemcpyAsync(d_taps, taps.data(), sizeof(int) * taps.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));  // `taps` is a local
      }
      CK(DMALLOC(&tmp, 4 * pn));
      dim3 gp(cdiv(W, 128), rh);
      k_conv_pass<<<gp, 128, 0, ctx->stream>>>(X + po, tmp, W, rh, F.r, F.kernel_kind, d_taps, total, 0); LAUNCHED();
      k_conv_pass<<<gp, 128, 0, ctx->stream>>>(tmp, X + po, W, rh, F.r, F.kernel_kind, d_taps, total, 1); LAUNCHED();
    }
  }
  // The geometry's matte in the update (render.ml:1099-1103).  Polygon.polygon_sprite samples every pixel it is
  // given, but a pixel whose 5 x 5 neighbourhood lies in the geometry's minshape has no edge anywhere near its
  // 2 x 2-pixel sampling window (a minshape pixel's row band [32y-47, 32y+16] and its columns are free of edge
  // pieces), so all 32 x 32 samples are inside and the opacity is 255: only the rest is super-sampled.
  CK(DMALLOC(&op, (size_t)nw * 32 * h)); CK(DMALLOC(&alpha, (size_t)W * h));
  {
    uint32_t* I = nullptr;   // interior = erode 2 2 minshape, clipped 2 pixels inside the rows / columns scanned here
    CK(DMALLOC(&I, 4 * nwords));
    k_bitop<<<wblocks, 256, 0, ctx->stream>>>(SG, CG, CG, nwords, 1); LAUNCHED();          // CG := minshape = shape - coverage
    k_fill_words<<<wblocks, 256, 0, ctx->stream>>>(I, nwords, 0xFFFFFFFFu); LAUNCHED();
    k_bitop<<<wblocks, 256, 0, ctx->stream>>>(I, CG, I, nwords, 1); LAUNCHED();            // everything but the minshape
    {  // rows y0 .. y1 only (everything outside is "not minshape" and the box mask below cuts 2 rows off each end)
      k_dilate<<<dim3(cdiv(nw, 128), h), 128, 0, ctx->stream>>>(I + (size_t)y0 * nw, R + (size_t)y0 * nw, h, nw, 2, 2); LAUNCHED();
    }
    k_fill_box_bits<<<dim3(cdiv(nw, 128), H), 128, 0, ctx->stream>>>(I, H, nw, 0, 0, 2, y0 + 2, W - 3, y1 - 2); LAUNCHED();
    k_bitop<<<wblocks, 256, 0, ctx->stream>>>(I, R, I, nwords, 1); LAUNCHED();             // interior
    k_bitop<<<wblocks, 256, 0, ctx->stream>>>(T, I, I, nwords, 1); LAUNCHED();             // to be super-sampled: T - interior
    CK(cudaMemsetAsync(op, 255, (size_t)nw * 32 * h, ctx->stream));
    k_aa_rows<<<dim3(cdiv(nw, 8), h), 256, 0, ctx->stream>>>(ed, F.count, F.winding, I + (size_t)y0 * nw, y0, h, 0, nw, ctx->d_aa, op, ctx->d_error); LAUNCHED();
    DFREE(I);
  }
  CK(cudaMemsetAsync(R, 0, 4 * nwords, ctx->stream));
  k_filter_matte<<<dim3(cdiv(nw, 4), h), 128, 0, ctx->stream>>>(T + (size_t)y0 * nw, op, F.colour, W, h, nw, alpha, R + (size_t)y0 * nw); LAUNCHED();  // R := finished
  k_bitop<<<wblocks, 256, 0, ctx->stream>>>(T, R, R, nwords, 1); LAUNCHED();      // pixels_for_normal_scene (render.ml:1105)
  CK(DMALLOC(&Z, 4 * (size_t)W * H));
  CK(cudaMemsetAsync(Z + (size_t)y0 * W, 0, 4 * (size_t)h * W, ctx->stream));
  if (render_suffix(ctx, s, F.pos, fi + 1, R, Z, true, tbox, true)) return 1;
  k_filter_blend<<<dim3(cdiv(W, 128), h), 128, 0, ctx->stream>>>(T + (size_t)y0 * nw, alpha, Z + (size_t)y0 * W, Y ? Y + (size_t)y0 * W : nullptr, target + (size_t)y0 * W, W, h, nw); LAUNCHED();
  k_bitop<<<wblocks, 256, 0, ctx->stream>>>(U, SG, U, nwords, 1); LAUNCHED();     // u --- ef (render.ml:1308)
  DFREE(SG); DFREE(CG); DFREE(T); DFREE(R); DFREE(X); DFREE(Z); DFREE(tmp); DFREE(op); DFREE(alpha); DFREE(d_taps);
  return 0;
}
// Render the scene list from leaf l0 / filter f0 to its end inside U (updated to the `u` left over)
static int render_suffix(coh_ctx* ctx, DevScene* s, int l0, int f0, uint32_t* U, uint32_t* target, bool fresh, PixBox box, bool target_zeroed) {
  const Frame& fr = ctx->fr;
  if (box.x1 < box.x0 || box.y1 < box.y0) return 0;
  auto segment = [&](int a, int b) -> int {
    if (b > a) {
      PassArgs A{a, b, box.x0, box.y0, box.x1 - box.x0 + 1, box.y1 - box.y0 + 1, U, U, target, fresh, !fresh};
      if (render_pass(ctx, s, A)) return 1;
    } else if (fresh && !target_zeroed) {
      const int hh = box.y1 - box.y0 + 1;
      k_clear_in_bits<<<dim3(cdiv(fr.W, 128), hh), 128, 0, ctx->stream>>>(target + (size_t)box.y0 * fr.W, U + (size_t)box.y0 * fr.tiles_x, fr.W, hh, fr.tiles_x); LAUNCHED();
    }
    fresh = false;
    return 0;
  };
  for (int f = f0; f < (int)s->filters.size(); f++) {
    if (segment(l0, s->filters[f].pos)) return 1;
    if (apply_filter(ctx, s, f, U, target, box)) return 1;
    l0 = s->filters[f].pos;
  }
  return segment(l0, s->n_scene_leaves);
}
static int render_filtered(coh_ctx* ctx, DevScene* s, const uint32_t* u_init, int ux, int uy, int uw, int uh) {
  const Frame& fr = ctx->fr;
  if (fr.band_y0 != 0 || fr.band_y1 != fr.H) FAIL("render_frame: scenes with filter objects need the whole frame on one context (filters read outside their band)");
  if (uw <= 0 || uh <= 0) return 0;
  const PixBox box{std::max(ux, 0), std::max(uy, 0), std::min(ux + uw - 1, fr.W - 1), std::min(uy + uh - 1, fr.H - 1)};
  if (box.x1 < box.x0 || box.y1 < box.y0) return 0;
  const int nw = fr.tiles_x;
  const size_t nwords = (size_t)nw * fr.H;
  uint32_t *U = ctx->u_out, *U0 = nullptr;
  CK(DMALLOC(&U0, 4 * nwords));
  if (u_init) CK(cudaMemcpyAsync(U0, u_init, 4 * nwords, cudaMemcpyDeviceToDevice, ctx->stream));
  else { k_fill_box_bits<<<dim3(cdiv(nw, 128), fr.H), 128, 0, ctx->stream>>>(U0, fr.H, nw, 0, 0, box.x0, box.y0, box.x1, box.y1); LAUNCHED(); }
  CK(cudaMemcpyAsync(U, U0, 4 * nwords, cudaMemcpyDeviceToDevice, ctx->stream));
  if (render_suffix(ctx, s, 0, 0, U, ctx->fb, true, box)) return 1;
  if (s->n_leaves > s->n_front_leaves) {
    // the background list shows wherever the scene pass is not opaque (render.ml:1363-1365)
    k_not_opaque_bits<<<dim3(cdiv(nw, 4), fr.H), 128, 0, ctx->stream>>>(ctx->fb, U0, U0, fr.W, fr.H, nw); LAUNCHED();
    PassArgs A{s->n_front_leaves, s->n_leaves, box.x0, box.y0, box.x1 - box.x0 + 1, box.y1 - box.y0 + 1, U0, nullptr, ctx->fb, false, true};
    if (render_pass(ctx, s, A)) return 1;
  }
  DFREE(U0);
  return 0;  // kernel-side failures are reported by coh_sync, as for plain frames
}

// Merge the objects of `scene` and `background` into one walk: the reference renders the two
// lists separately over the same update and composites the results with `over`
// (render.ml:1357-1365); a pixel of the background is only visible where the scene pass left
// `u`, so one front-to-back walk over [Group scene; Group background] gives the same pixels.
int coh_render_frame(coh_ctx* ctx, coh_scene_t scene, int32_t ux, int32_t uy, int32_t uw, int32_t uh, int32_t flags) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fb) FAIL("coh_render_frame: call coh_fb_configure first");
  if (uw < 0 || uh < 0) FAIL("Sprite.box: negative argument.");
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_render_frame: null scene");
  bool record_u = (flags & COH_RENDER_RECORD_U) != 0;
  if (!s->filters.empty()) {
    if (render_filtered(ctx, s, nullptr, ux, uy, uw, uh)) return 1;
    ctx->have_u = true;
    return 0;
  }
  PassArgs A{0, s->n_leaves, ux, uy, uw, uh, nullptr, record_u ? ctx->u_out : nullptr, ctx->fb, true, false};
  if (render_pass(ctx, s, A)) return 1;
  ctx->have_u = record_u;
  return 0;
}
// ---------------------------------------------------------------------------------------
// Cache (cache.mli:32-48): span sets resident in HBM, keyed by id
// ---------------------------------------------------------------------------------------
static size_t shape_bytes(const DevShape* s) { return s ? sizeof(int) * (s->n_rows + 1) + sizeof(int2) * (size_t)s->n_spans : 0; }
static DevShape* clone_shape(coh_ctx* ctx, const DevShape* s, int dx, int dy) {
  if (!s) return nullptr;
  coh_shape_t out = 0;
  if (coh_shape_translate(ctx, (coh_shape_t)s, dx, dy, &out)) return nullptr;
  return (DevShape*)out;
}
static void cache_drop(coh_ctx* ctx, std::map<int64_t, CacheEntry>::iterator it) {
  ctx->cache_size -= it->second.bytes;
  free_shape(ctx, it->second.shape); free_shape(ctx, it->second.minshape);
  ctx->cache.erase(it);
}
static void cache_drophalf(coh_ctx* ctx) {  // cache.ml:242-271 (eviction order: least recently used first)
  size_t target = ctx->cache_size / 2;
  while (ctx->cache_size > target) {
    auto victim = ctx->cache.end();
    for (auto it = ctx->cache.begin(); it != ctx->cache.end(); ++it)
      if (!it->second.alias && it->second.has && (victim == ctx->cache.end() || it->second.lastused < victim->second.lastused)) victim = it;
    if (victim == ctx->cache.end()) break;
    const int64_t vid = victim->first;
    cache_drop(ctx, victim);
    for (auto it = ctx->cache.begin(); it != ctx->cache.end();)  // aliases go with their parent (cache.ml:119-127)
      if (it->second.alias && it->second.target == vid) it = ctx->cache.erase(it); else ++it;
  }
}
int coh_cache_clear(coh_ctx* ctx) {
  CK(cudaSetDevice(ctx->device));
  while (!ctx->cache.empty()) cache_drop(ctx, ctx->cache.begin());
  ctx->cache_size = 0;
  return 0;
}
int coh_cache_configure(coh_ctx* ctx, int32_t usecache, int64_t max_bytes) {  // Cache.usecache, Cache.setsize
  ctx->usecache = usecache != 0;
  if (max_bytes > 0) { ctx->cache_max = (size_t)max_bytes; while (ctx->cache_size > ctx->cache_max) cache_drophalf(ctx); }
  return 0;
}
int coh_cache_stats(coh_ctx* ctx, int64_t out[4]) {  // cache.ml:24-38
  out[0] = ctx->shphit; out[1] = ctx->shpmis; out[2] = (int64_t)ctx->cache_size; out[3] = (int64_t)ctx->cache.size();
  return 0;
}
// Cache.addshape idset shp minshp (cache.ml:280-324): copies are kept; an existing shape is not replaced
int coh_cache_addshape(coh_ctx* ctx, int64_t id, coh_shape_t shape, coh_shape_t minshape) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->usecache || id < 0) return 0;
  size_t bytes = shape_bytes((DevShape*)shape) + shape_bytes((DevShape*)minshape);
  if (bytes > ctx->cache_max / 2) return 0;
  if (ctx->cache_size + bytes > ctx->cache_max) cache_drophalf(ctx);
  auto it = ctx->cache.find(id);
  int dx = 0, dy = 0;
  if (it != ctx->cache.end() && it->second.alias) { dx = it->second.dx; dy = it->second.dy; id = it->second.target; it = ctx->cache.find(id); }
  if (it != ctx->cache.end() && it->second.has) return 0;
  CacheEntry& e = ctx->cache[id];
  e.shape = clone_shape(ctx, (DevShape*)shape, -dx, -dy); e.minshape = clone_shape(ctx, (DevShape*)minshape, -dx, -dy);
  e.has = true; e.bytes = bytes; e.lastused = ++ctx->cache_timer;
  ctx->cache_size += bytes;
  return 0;
}
// Cache.getshape idset (cache.ml:370-387): fresh handles (translated through aliases); found = 0 on a miss
int coh_cache_getshape(coh_ctx* ctx, int64_t id, coh_shape_t* shape, coh_shape_t* minshape, int32_t* found) {
  CK(cudaSetDevice(ctx->device));
  *shape = 0; *minshape = 0; *found = 0;
  if (!ctx->usecache || id < 0) return 0;
  auto it = ctx->cache.find(id);
  int dx = 0, dy = 0;
  if (it != ctx->cache.end() && it->second.alias) { dx = it->second.dx; dy = it->second.dy; it = ctx->cache.find(it->second.target); }
  if (it == ctx->cache.end() || !it->second.has) { ctx->shpmis++; return 0; }
  ctx->shphit++; it->second.lastused = ++ctx->cache_timer;
  *shape = (coh_shape_t)clone_shape(ctx, it->second.shape, dx, dy);
  *minshape = (coh_shape_t)clone_shape(ctx, it->second.minshape, dx, dy);
  *found = 1;
  return 0;
}
// Cache.addtranslation idset target dx dy (cache.ml:423-436)
int coh_cache_addtranslation(coh_ctx* ctx, int64_t id, int64_t target, int32_t dx, int32_t dy) {
  if (!ctx->usecache) return 0;
  ctx->cache_timer++;
  auto it = ctx->cache.find(target);
  if (it == ctx->cache.end()) return 0;  // not in the cache, so can't add a translation
  CacheEntry e; e.alias = true;
  if (it->second.alias) { e.dx = dx + it->second.dx; e.dy = dy + it->second.dy; e.target = it->second.target; }
  else { e.dx = dx; e.dy = dy; e.target = target; }
  ctx->cache[id] = e;
  return 0;
}
// Render.shape_of_basicshape obj (render.ml:469-594) for the obj_index-th object of a scene, through
// the cache: the entry is keyed by the object's id and holds the shape of the UNTRANSLATED geometry;
// the object's alias offset is applied on the way out (cache.ml:380-385).
static int object_shape_rec(coh_ctx* ctx, DevScene* s, int r, coh_shape_t* shape, coh_shape_t* minshape) {
  const ObjRec& o = s->h_objs[r];
  *shape = 0; *minshape = 0;
  const int64_t id = s->ids[r];
  int32_t found = 0;
  coh_shape_t cs = 0, cm = 0;
  if (o.kind != K_GROUP && coh_cache_getshape(ctx, id, &cs, &cm, &found)) return 1;   // group shapes are kept per scene
  if (!found) {
    if (o.kind == K_PATH) {
      if (shapes_from_device_edges(ctx, s->edges + o.first, o.count, o.winding, o.bx0 - o.dx, o.by0 - o.dy, o.bx1 - o.dx, o.by1 - o.dy, &cs, &cm, "coh_scene_object_shape")) return 1;
    } else if (o.kind == K_CPG) {  // render.ml:508-528
      coh_shape_t as = 0, am = 0, bs = 0, bm = 0, t0 = 0, t1 = 0;
      const int x0 = o.bx0 - o.dx, y0 = o.by0 - o.dy, x1 = o.bx1 - o.dx, y1 = o.by1 - o.dy;
      if (o.count && shapes_from_device_edges(ctx, s->edges + o.first, o.count, o.winding, x0, y0, x1, y1, &as, &am, "coh_scene_object_shape")) return 1;
      if (o.b_count && shapes_from_device_edges(ctx, s->edges + o.b_first, o.b_count, o.b_opw >> 8, x0, y0, x1, y1, &bs, &bm, "coh_scene_object_shape")) return 1;
      int rc = 0;
      switch (o.b_opw & 255) {
        case COH_CPG_UNION: rc = coh_shape_union(ctx, as, bs, &cs) || coh_shape_union(ctx, am, bm, &cm); break;
        case COH_CPG_INTERSECTION: rc = coh_shape_intersection(ctx, as, bs, &cs) || coh_shape_intersection(ctx, am, bm, &cm); break;
        case COH_CPG_SUBTRACTION: rc = coh_shape_difference(ctx, as, bm, &cs) || coh_shape_difference(ctx, am, bs, &cm); break;
        default:
          rc = coh_shape_union(ctx, as, bs, &t0) || coh_shape_intersection(ctx, am, bm, &t1) || coh_shape_difference(ctx, t0, t1, &cs);
          coh_shape_free(ctx, t0); coh_shape_free(ctx, t1); t0 = t1 = 0;
          rc = rc || coh_shape_difference(ctx, bm, as, &t0) || coh_shape_difference(ctx, am, bs, &t1) || coh_shape_union(ctx, t0, t1, &cm);
          coh_shape_free(ctx, t0); coh_shape_free(ctx, t1);
      }
      coh_shape_free(ctx, as); coh_shape_free(ctx, am); coh_shape_free(ctx, bs); coh_shape_free(ctx, bm);
      if (rc) return 1;
    } else if (o.kind == K_PRIM) {
      if (coh_shape_box(ctx, o.prim[0], o.prim[1], o.prim[2] - o.prim[0] + 1, o.prim[3] - o.prim[1] + 1, &cs)) return 1;
      if (coh_shape_translate(ctx, cs, 0, 0, &cm)) return 1;
    } else if (o.kind == K_GROUP) {
      // union of the members' shapes, minshape null (render.ml:476-496); members are not cached (fresh ids)
      const int2 off = s->group_off[r];
      auto git = ctx->usecache ? s->group_shape.find(r) : s->group_shape.end();
      if (git != s->group_shape.end()) {
        ctx->shphit++;
        return coh_shape_translate(ctx, (coh_shape_t)git->second.shape, off.x - git->second.offx, off.y - git->second.offy, shape);
      }
      for (int k = r + 1; k <= s->group_last[r]; k++) {
        if (s->h_objs[k].depth != o.depth + 1) continue;  // direct children only (nested groups recurse)
        coh_shape_t ms = 0, mm = 0, un = 0;
        if (object_shape_rec(ctx, s, k, &ms, &mm)) return 1;
        if (coh_shape_union(ctx, cs, ms, &un)) return 1;
        coh_shape_free(ctx, cs); coh_shape_free(ctx, ms); coh_shape_free(ctx, mm);
        cs = un;
      }
      // members already carry their own alias offsets
      if (ctx->usecache && cs) {
        coh_shape_t keep = 0;
        if (coh_shape_translate(ctx, cs, 0, 0, &keep)) return 1;
        s->group_shape[r] = DevScene::GroupShape{(DevShape*)keep, off.x, off.y};
      }
      *shape = cs; *minshape = 0;
      return 0;
    } else if (o.kind == K_BRUSH) {  // Brush.shape_of_brushstroke, minshape null (render.ml:529-535)
      const int x0 = o.bx0 - o.dx, y0 = o.by0 - o.dy, x1 = o.bx1 - o.dx, y1 = o.by1 - o.dy;
      const int wx0 = floordiv(x0, 32) * 32, nw = (x1 - wx0) / 32 + 1, n_rows = y1 - y0 + 1;
      uint32_t* bits = nullptr;
      CK(DMALLOC(&bits, 4 * (size_t)nw * n_rows));
      CK(cudaMemsetAsync(bits, 0, 4 * (size_t)nw * n_rows, ctx->stream));
      const int side = 2 * o.brush_r + 1;
      k_stamp_boxes_to_bits<<<cdiv(o.count * side, 256), 256, 0, ctx->stream>>>(s->points + o.first, o.count, o.brush_r, y0, n_rows, wx0, nw, bits); LAUNCHED();
      int rc = shape_from_bits(ctx, bits, y0, n_rows, wx0, nw, &cs);
      DFREE(bits);
      if (rc) return 1;
    } else if (o.kind == K_CONV) {   // bloat r r (shape g), erode r r (minshape g) (render.ml:536-555): kept as bit-rows by the scene
      const uint32_t* S = s->conv_bits + o.cv_bits;
      if (shape_from_bits(ctx, S, o.cv_y0, o.cv_h, o.cv_x0, o.cv_nw, &cs)) return 1;
      if (shape_from_bits(ctx, S + (size_t)o.cv_nw * o.cv_h, o.cv_y0, o.cv_h, o.cv_x0, o.cv_nw, &cm)) return 1;
    } else FAIL("coh_scene_object_shape: unsupported object kind");
    if (coh_cache_addshape(ctx, id, cs, cm)) return 1;
  }
  // apply the alias offset
  if (o.dx || o.dy) {
    coh_shape_t ts = 0, tm = 0;
    if (coh_shape_translate(ctx, cs, o.dx, o.dy, &ts) || coh_shape_translate(ctx, cm, o.dx, o.dy, &tm)) return 1;
    coh_shape_free(ctx, cs); coh_shape_free(ctx, cm);
    cs = ts; cm = tm;
  }
  *shape = cs; *minshape = cm;
  return 0;
}
int coh_scene_object_shape(coh_ctx* ctx, coh_scene_t scene, int32_t obj_index, coh_shape_t* shape, coh_shape_t* minshape) {
  CK(cudaSetDevice(ctx->device));
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_scene_object_shape: null scene");
  if (obj_index < 0 || obj_index >= (int)s->rec_of_abi.size() || s->rec_of_abi[obj_index] < 0) FAIL("coh_scene_object_shape: no such object");
  return object_shape_rec(ctx, s, s->rec_of_abi[obj_index], shape, minshape);
}
// Render.plaindirty / alldirty (render.ml:1376-1391): ((shp_o - minshp_n) ∪ (shp_n - minshp_o)) ∩ u,
// or (shp_o ∪ shp_n) ∩ u when `plain` is 0.
int coh_dirty_region(coh_ctx* ctx, coh_shape_t shp_o, coh_shape_t min_o, coh_shape_t shp_n, coh_shape_t min_n,
                     coh_shape_t u, int32_t plain, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  coh_shape_t a = 0, b = 0, c = 0;
  if (plain) {
    if (coh_shape_difference(ctx, shp_o, min_n, &a) || coh_shape_difference(ctx, shp_n, min_o, &b)) return 1;
    if (coh_shape_union(ctx, a, b, &c)) return 1;
    coh_shape_free(ctx, a); coh_shape_free(ctx, b);
  } else {
    if (coh_shape_union(ctx, shp_o, shp_n, &c)) return 1;
  }
  int rc = coh_shape_intersection(ctx, c, u, out);
  coh_shape_free(ctx, c);
  return rc;
}

// Render.translate_renderobject dx dy obj (render.ml:259-271): the object (or every member of the
// group) becomes an alias of its former self moved by whole pixels; only the alias offsets and the
// boxes the binning reads change, nothing is re-uploaded.
int coh_scene_translate_object(coh_ctx* ctx, coh_scene_t scene, int32_t obj_index, int32_t dx, int32_t dy) {
  CK(cudaSetDevice(ctx->device));
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_scene_translate_object: null scene");
  if (obj_index < 0 || obj_index >= (int)s->rec_of_abi.size() || s->rec_of_abi[obj_index] < 0) FAIL("coh_scene_translate_object: no such object");
  const int r = s->rec_of_abi[obj_index];
  const int last = s->h_objs[r].kind == K_GROUP ? s->group_last[r] : r;
  if (s->h_objs[r].kind != K_GROUP)   // a member moved on its own: the shapes of the groups around it are stale
    for (int d = 0; d < s->h_objs[r].depth; d++) {
      auto git = s->group_shape.find(s->h_objs[r].anc[d]);
      if (git != s->group_shape.end()) { free_shape(ctx, git->second.shape); s->group_shape.erase(git); }
    }
  for (int k = r; k <= last; k++) {
    ObjRec& o = s->h_objs[k];
    if (o.kind == K_GROUP) { s->group_off[k].x += dx; s->group_off[k].y += dy; continue; }
    o.dx += dx; o.dy += dy; o.bx0 += dx; o.bx1 += dx; o.by0 += dy; o.by1 += dy;
  }
  s->items_for_W = -1;  // the item-pool bound depends on the boxes
  if (s->n_leaves > 0) { k_move_leaves<<<cdiv(s->n_leaves, 256), 256, 0, ctx->stream>>>(s->objs, s->leaf_box, s->leaves, s->n_leaves, r, last, dx, dy); LAUNCHED(); }
  return 0;
}
// Render.render_frame over an arbitrary update shape (the dirty region of engine.ml:224-252).
int coh_render_frame_shape(coh_ctx* ctx, coh_scene_t scene, coh_shape_t update, int32_t flags) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fb) FAIL("coh_render_frame_shape: call coh_fb_configure first");
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_render_frame_shape: null scene");
  DevShape* us = (DevShape*)update;
  ctx->have_u = false;
  if (!us) return 0;  // NullShape: nothing to render (render.ml:1321-1322)
  const Frame& fr = ctx->fr;
  if (!ctx->u_init) CK(DMALLOC(&ctx->u_init, sizeof(uint32_t) * (size_t)fr.tiles_x * fr.H));
  CK(cudaMemsetAsync(ctx->u_init, 0, sizeof(uint32_t) * (size_t)fr.tiles_x * fr.H, ctx->stream));
  k_spans_to_bits<<<cdiv(fr.H, 128), 128, 0, ctx->stream>>>(us->row_ptr, us->spans, us->y0, us->n_rows, 0, fr.H, 0, fr.tiles_x, ctx->u_init); LAUNCHED();
  bool record_u = (flags & COH_RENDER_RECORD_U) != 0;
  if (!s->filters.empty()) {
    int rc = render_filtered(ctx, s, ctx->u_init, us->bx0, us->by0, us->bx1 - us->bx0 + 1, us->by1 - us->by0 + 1);
    ctx->have_u = !rc;
    return rc;
  }
  PassArgs A{0, s->n_leaves, us->bx0, us->by0, us->bx1 - us->bx0 + 1, us->by1 - us->by0 + 1, ctx->u_init, record_u ? ctx->u_out : nullptr, ctx->fb, true, false};
  int rc = render_pass(ctx, s, A);
  ctx->have_u = record_u && !rc;
  return rc;
}
// Render.dirty_filter (render.ml:1418-1438) with the dirty functions of filters.ml restated per filter kind.
int coh_dirty_filter(coh_ctx* ctx, coh_scene_t scene, int32_t lmo_index, coh_shape_t initial_dirty, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_dirty_filter: null scene");
  coh_shape_t cur = 0;
  if (coh_shape_translate(ctx, initial_dirty, 0, 0, &cur)) return 1;
  // filters above the lmo, folded from the last of them to the first (fold_left over rev filters)
  for (int k = (int)s->filters.size() - 1; k >= 0; k--) {
    const DevScene::FilterRec& F = s->filters[k];
    if (lmo_index >= 0 && F.abi >= lmo_index) continue;
    if (F.kind != COH_FILTER_BLUR || !cur) continue;  // nulldirty
    // bloatdirty r r (filters.ml:63-75)
    coh_shape_t fs = 0, fm = 0, bf = 0, inf = 0, outf = 0, bl = 0, bif = 0, res = 0;
    if (shapes_from_device_edges(ctx, s->edges + F.first, F.count, F.winding, F.bx0, F.by0, F.bx1, F.by1, &fs, &fm, "coh_dirty_filter")) return 1;
    int rc = coh_shape_bloat(ctx, fs, F.r, F.r, &bf) || coh_shape_intersection(ctx, bf, cur, &inf) || coh_shape_difference(ctx, cur, bf, &outf) ||
             coh_shape_bloat(ctx, inf, F.r, F.r, &bl) || coh_shape_intersection(ctx, bl, bf, &bif) || coh_shape_union(ctx, bif, outf, &res);
    for (coh_shape_t h : {fs, fm, bf, inf, outf, bl, bif, cur}) coh_shape_free(ctx, h);
    if (rc) return 1;
    cur = res;
  }
  *out = cur;
  return 0;
}
// One drag step on device-resident data (see the header): translate, dirty region as a bit-frame, render.
int coh_scene_drag_object(coh_ctx* ctx, coh_scene_t scene, int32_t obj_index, int32_t dx, int32_t dy, int32_t flags,
                          int32_t dirty_bbox[4]) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fb) FAIL("coh_scene_drag_object: call coh_fb_configure first");
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_scene_drag_object: null scene");
  if (obj_index < 0 || obj_index >= (int)s->rec_of_abi.size() || s->rec_of_abi[obj_index] < 0) FAIL("coh_scene_drag_object: no such object");
  const int r = s->rec_of_abi[obj_index];
  const ObjRec& o = s->h_objs[r];
  // Fill.Plain objects: plaindirty; groups, fancy fills, brush strokes, Convolved objects: alldirty (render.ml:1396-1400)
  const bool plain = (o.kind == K_PATH || o.kind == K_CPG) && o.fill.kind == 0;
  const bool prim = o.kind == K_PRIM;
  coh_shape_t so = 0, mo = 0;
  if (object_shape_rec(ctx, s, r, &so, &mo)) return 1;   // served by the cache after the first step
  if (coh_scene_translate_object(ctx, scene, obj_index, dx, dy)) return 1;
  const Frame& fr = ctx->fr;
  const int nw = fr.tiles_x;
  const size_t nwords = (size_t)nw * fr.H;
  if (!ctx->u_init) CK(DMALLOC(&ctx->u_init, sizeof(uint32_t) * nwords));
  uint32_t* U = ctx->u_init;
  CK(cudaMemsetAsync(U, 0, 4 * nwords, ctx->stream));
  DevShape* S = (DevShape*)so; DevShape* M = (DevShape*)mo;
  int bb[4] = {0, 0, -1, -1};
  if (S) {
    // old position: offset 0; new position: the same span set read through the offset (dx, dy)
    auto put = [&](const DevShape* sh, uint32_t* bits, int ox, int oy) -> int {
      k_spans_to_bits<<<cdiv(fr.H, 128), 128, 0, ctx->stream>>>(sh->row_ptr, sh->spans, sh->y0 + oy, sh->n_rows, 0, fr.H, -ox, nw, bits); LAUNCHED();
      return 0;
    };
    if ((plain || prim) && M) {
      uint32_t *A = nullptr, *B = nullptr;
      CK(DMALLOC(&A, 4 * nwords)); CK(DMALLOC(&B, 4 * nwords));
      const unsigned wb = (unsigned)((nwords + 255) / 256);
      CK(cudaMemsetAsync(A, 0, 4 * nwords, ctx->stream)); CK(cudaMemsetAsync(B, 0, 4 * nwords, ctx->stream));
      if (put(S, A, 0, 0) || put(M, B, dx, dy)) return 1;
      k_bitop<<<wb, 256, 0, ctx->stream>>>(A, B, U, nwords, 1); LAUNCHED();            // shp_o --- minshp_n
      CK(cudaMemsetAsync(A, 0, 4 * nwords, ctx->stream)); CK(cudaMemsetAsync(B, 0, 4 * nwords, ctx->stream));
      if (put(S, A, dx, dy) || put(M, B, 0, 0)) return 1;
      k_bitop<<<wb, 256, 0, ctx->stream>>>(A, B, A, nwords, 1); LAUNCHED();            // shp_n --- minshp_o
      k_bitop<<<wb, 256, 0, ctx->stream>>>(U, A, U, nwords, 0); LAUNCHED();
      DFREE(A); DFREE(B);
    } else {
      if (put(S, U, 0, 0) || put(S, U, dx, dy)) return 1;                                // shp_o ||| shp_n
    }
    bb[0] = std::max(0, S->bx0 + std::min(dx, 0)); bb[1] = std::max(0, S->by0 + std::min(dy, 0));
    bb[2] = std::min(fr.W - 1, S->bx1 + std::max(dx, 0)); bb[3] = std::min(fr.H - 1, S->by1 + std::max(dy, 0));
  }
  coh_shape_free(ctx, so); coh_shape_free(ctx, mo);
  if (dirty_bbox) for (int k = 0; k < 4; k++) dirty_bbox[k] = bb[k];
  ctx->have_u = false;
  if (bb[2] < bb[0] || bb[3] < bb[1]) return 0;
  const bool record_u = (flags & COH_RENDER_RECORD_U) != 0;
  int rc;
  if (!s->filters.empty()) { rc = render_filtered(ctx, s, U, bb[0], bb[1], bb[2] - bb[0] + 1, bb[3] - bb[1] + 1); ctx->have_u = !rc; return rc; }
  PassArgs A{0, s->n_leaves, bb[0], bb[1], bb[2] - bb[0] + 1, bb[3] - bb[1] + 1, U, record_u ? ctx->u_out : nullptr, ctx->fb, true, false};
  rc = render_pass(ctx, s, A);
  ctx->have_u = record_u && !rc;
  return rc;
}
int coh_render_uncovered(coh_ctx* ctx, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (!ctx->have_u) FAIL("coh_render_uncovered: no frame rendered");
  const Frame& fr = ctx->fr;
  int n_rows = fr.band_y1 - fr.band_y0;
  return shape_from_bits(ctx, ctx->u_out + (size_t)fr.band_y0 * fr.tiles_x, fr.band_y0, n_rows, 0, fr.tiles_x, out);
}
void* coh_fb_device_ptr(coh_ctx* ctx) { return ctx->fb; }
#ifdef COH_PHASE_PROFILE
int coh_cell_cycles(coh_ctx* ctx, unsigned int* out, int n) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpyFromSymbol(out, g_cell_cycles, sizeof(unsigned int) * n));
  return 0;
}
int coh_phase_cycles(coh_ctx* ctx, unsigned long long* out8, int reset) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpyFromSymbol(out8, g_phase_cycles, sizeof(unsigned long long) * 16));
  if (reset) { unsigned long long z[16] = {0}; CK(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z)); }
  return 0;
}
#endif
int coh_fb_read_rgba(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fb) FAIL("coh_fb_read_rgba: no framebuffer");
  if (x < 0 || y < 0 || w < 0 || h < 0 || x + w > ctx->fr.W || y + h > ctx->fr.H) FAIL("coh_fb_read_rgba: rectangle outside the framebuffer");
  if (w == 0 || h == 0) return 0;
  CK(cudaMemcpy2DAsync(out, (size_t)w * 4, ctx->fb + (size_t)y * ctx->fr.W + x, (size_t)ctx->fr.W * 4, (size_t)w * 4, h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int coh_fb_read_rgba_async(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fb) FAIL("coh_fb_read_rgba_async: no framebuffer");
  if (x < 0 || y < 0 || w < 0 || h < 0 || x + w > ctx->fr.W || y + h > ctx->fr.H) FAIL("coh_fb_read_rgba_async: rectangle outside the framebuffer");
  if (w == 0 || h == 0) return 0;
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) { CK(cudaEventCreateWithFlags(&ctx->ev_ready[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming)); }
  }
  const int k = ctx->stage_next; ctx->stage_next ^= 1;
  const size_t bytes = (size_t)w * h * 4;
  if (bytes > ctx->stage_cap[k]) {
    if (ctx->stage_busy[k]) CK(cudaEventSynchronize(ctx->ev_done[k]));
    DFREE(ctx->stage[k]);
    CK(DMALLOC(&ctx->stage[k], bytes));
    ctx->stage_cap[k] = bytes;
  }
  if (ctx->stage_busy[k]) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_done[k], 0));  // the copy of two reads ago still owns this buffer
  CK(cudaMemcpy2DAsync(ctx->stage[k], (size_t)w * 4, ctx->fb + (size_t)y * ctx->fr.W + x, (size_t)ctx->fr.W * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev_ready[k], ctx->stream));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready[k], 0));
  CK(cudaMemcpyAsync(out, ctx->stage[k], bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaEventRecord(ctx->ev_done[k], ctx->copy_stream));
  ctx->stage_busy[k] = true;
  return 0;
}
int coh_fb_read_wait(coh_ctx* ctx) {
  CK(cudaSetDevice(ctx->device));
  for (int k = 0; k < 2; k++)
    if (ctx->stage_busy[k]) { CK(cudaEventSynchronize(ctx->ev_done[k])); ctx->stage_busy[k] = false; }
  return 0;
}
int coh_fb_read_rgb888(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fb) FAIL("coh_fb_read_rgb888: no framebuffer");
  if (x < 0 || y < 0 || w < 0 || h < 0 || x + w > ctx->fr.W || y + h > ctx->fr.H) FAIL("coh_fb_read_rgb888: rectangle outside the framebuffer");
  if (w == 0 || h == 0) return 0;
  uint8_t* d = nullptr;
  CK(DMALLOC(&d, (size_t)w * h * 3));
  dim3 g(cdiv(w, 128), h);
  k_rgb888<<<g, 128, 0, ctx->stream>>>(ctx->fb, ctx->fr.W, x, y, w, h, d); LAUNCHED();
  CK(cudaMemcpyAsync(out, d, (size_t)w * h * 3, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(d);
  return 0;
}

}  // extern "C"
