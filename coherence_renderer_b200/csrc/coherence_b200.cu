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
#include <chrono>
#include <memory>
#include <atomic>
#include <thread>
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
// Output of K1 cell binning for one pass geometry: per-cell front-to-back lists and the heavy-first order.
struct BinSet {
  int n_cells_cap = 0;
  int2* cell_head = nullptr;  // per cell: colour + flags when the cell is one opaque covering primitive
  int2* cell_rng = nullptr;   // per cell [start, end) into cell_items
  int* cell_order = nullptr;  // cells by list-length class [BIN_CLASSES][n_cells]
  bool comp_valid = false;
  int4* comp_order = nullptr; // the same order, flattened for the row compositor: {cell, list start, list end, header flags}, cell -1 past the end
  int* cell_items = nullptr; size_t cell_items_cap = 0;
  int* item_cell = nullptr;   // cell of every list entry (small-scene binning only)
  int2* item_attr = nullptr;  // compositing attributes of every list entry (small-scene binning only)
  int4* item_rec = nullptr;   // scan-conversion record of every list entry, two int4 each (small-scene binning only)
  int* state = nullptr;       // ORDER_BINS ints: [0] pool cursor, [1 ..] class counts
};
// A leaf list of a scene (front to back) with what the binning derives from it.  A scene has two: every leaf, and
// the list in which every object with a cached sprite is ONE leaf (its members collapsed).
struct LeafView {
  int n = 0;
  int* leaves = nullptr;
  int4* leaf_box = nullptr;     // conservative device-space pixel box per leaf (binning reads these, coalesced)
  std::vector<int> h_leaves;
  size_t items_total = 0, coarse_total = 0; bool coarse_total_valid = false;
  int items_for_W = -1, items_for_H = -1, items_for_y0 = -1, items_for_y1 = -1;
  // Whole-frame binning kept with the scene (it is a pure function of the object boxes and the frame geometry,
  // like the row-edge lists): valid until an object moves or the framebuffer geometry changes.
  BinSet bins; bool bins_valid = false; int bins_key[6] = {0, 0, 0, 0, 0, 0};
};
// Partial-sprite cache entry (render.ml:1169-1242; cache.ml:328-367): a top-level Group with an id, as one sprite.
struct SpriteEntry {
  int grp = 0;                 // record of the group
  int l0 = 0, l1 = 0;          // its member leaves in the full list
  int leaf_rec = 0;            // record of the sprite leaf (K_CONV layout: shape planes + RGBA8 canvas in the object's frame)
  uint32_t* valid = nullptr;   // pshape: pixels whose value is in the canvas
  size_t plane_words = 0, bytes = 0;
  bool complete = false;       // pshape = shape: nothing left to render, ever
  bool dead = false;           // a member moved on its own: the entry is not used any more
  int* h_missing = nullptr;    // pinned: pixels of the shape not yet in pshape (read back asynchronously)
  int* d_missing = nullptr;
  cudaEvent_t ev = nullptr; bool ev_pending = false;
};
struct DevScene {
  int n_objs = 0, n_edges = 0, n_points = 0;
  ObjRec* objs = nullptr;
  LeafView full, sp;            // every leaf / cached objects collapsed into their sprite leaves
  int& n_leaves = full.n;
  int*& leaves = full.leaves;
  int4*& leaf_box = full.leaf_box;
  std::vector<SpriteEntry> sprites;
  int64_t sprite_hits = 0, sprite_fills = 0;
  EdgeRec* edges = nullptr;
  int2* points = nullptr;
  uint8_t* stamps = nullptr;
  int* rowedge_ptr = nullptr;   // K1 edge binning (CSR over (path object, pixel row))
  int* rowedge_idx = nullptr;
  int2* brush_ranges = nullptr; // per (stroke, cell of its box): [first, last] stamp index reaching the cell
  uint32_t* conv_bits = nullptr; // Convolved objects: shape / minshape bit-rows
  uint32_t* conv_px = nullptr;   // Convolved objects: pre-convolved canvases
  int2* attr = nullptr;          // per record: {plain colour, is-path | background-list << 1 | occludes << 2 | (pretrans + 1) << 8}
  bool flat_ok = false;          // every leaf a direct member of a root list, plain paths / primitives only
  std::vector<ObjRec> h_objs;
  std::vector<int64_t> ids;      // cache key (Id.idset) of every record
  std::vector<int> rec_of_abi;   // record index of every object of the ABI array (-1: GROUP_END / dropped)
  std::vector<int> group_last;   // for group records: last record index inside the group
  std::vector<int> real_depth;   // per record: enclosing groups in the scene as given (ObjRec.depth leaves out groups dissolved into their parent)
  // Filters (render.ml:37-48): top-level members of the scene list that are not leaves.  `pos` = number of
  // ordinary scene leaves in front of the filter; the leaves are ordered [scene | reading scenes | background].
  struct FilterRec { int pos, kind, kernel_kind, r, first, count, winding, aa_winding; uint32_t colour; int read0, read1; int bx0, by0, bx1, by1; int abi; int taps_off, taps_total;
                     int dx, dy;   // alias translation in whole pixels (render.ml:259-271); bx0 .. by1 include it
                     int head_abi, head_l1;   // MINUS: the object that follows the filter, the leaf index where the list continues after it
                     DevScene* geom_sub; int gcx0, gcy0, gcnw, gch;   // COH_GEOM_NEXT: the geometry object as a scene of its own, aliased into a canvas at (gcx0, gcy0) of gcnw words x gch rows
                     int first2, count2, stamp_off, brush_r;   // SMEAR: smear points (in the points array), the brush's stamp (alpha bytes), its radius
                     // kept with the scene once computed (the reference finds a filter geometry's shape in its cache by id,
                     // render.ml:472-474): shape / coverage bit-rows and the antialiased opacity of every shape pixel, for the
                     // geometry's rows gy0 .. gy0 + gh - 1 of a frame gW x gH
                     uint32_t *SG, *CG; uint8_t* op; int gy0, gh, gW, gH, gdx, gdy; };
  int* filter_taps = nullptr;    // blur filters: Convolve.mkgaussian taps (convolve.ml:60-70), made once per scene
  std::vector<FilterRec> filters;
  // Group shapes (render.ml:476-496 caches them under the group's id): kept per scene, in the frame the group
  // had when the entry was made; moving the whole group only changes the offset applied on the way out.
  struct GroupShape { DevShape* shape; int offx, offy; };
  std::map<int, GroupShape> group_shape;
  std::vector<int2> group_off;   // per record: translation applied to the whole group since scene creation
  int n_scene_leaves = 0;        // ordinary leaves of the scene list
  int n_front_leaves = 0;        // + leaves of reading-scene groups (the background list follows)
  std::vector<int>& h_leaves = full.h_leaves;
  bool has_fancy = false;    // some object has a gradient / radial fill
  int extras = 0;            // walker variant: 0 polygons / primitives, 1 + brush / Convolved, 2 + CPG / filters
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
  uint32_t* touched = nullptr; // while a smear filter renders its reading scene: the bit-frame that receives the shape of that render
  bool use_u_init = false;
  bool have_u = false;
  // binning scratch (passes whose binning is not kept with the scene)
  BinSet bins;
  int opt_pre_min_pairs_walk = 1 << 30;  // the same threshold for passes that composite with the walker when their cells are few (one-row items)
  int opt_pre_min_pairs = 4096;  // passes with fewer (list entry, row) pairs stay on the fused walker: four dependent launches cost more than they gain
  bool opt_bin_cache = true;  // keep whole-frame binning with the scene
  bool opt_fork_prefill = true;  // three-phase frames: background prefill on a second stream beside the scan kernels
  int opt_ab = 0;   // scratch switch for A/B measurements
  bool opt_comp_rows = true;  // flat scenes: row compositor instead of the walker in three-phase frames
  bool opt_pdl = true;  // three-phase frames: the chain of kernels launched with programmatic stream serialization
  // large scenes: coarse level of the two-level binning (leaf positions per coarse cell)
  int* coarse_items = nullptr; int* coarse_counts = nullptr; int* coarse_off = nullptr; size_t coarse_cap = 0, coarse_cells_cap = 0;
  uint32_t* peer_fb[COH_MAX_PEERS] = {nullptr}; int n_peers = 0;  // coh_fb_set_peers
  // three-phase frames: per (cell item, row) pair
  uint2* pre_sc = nullptr; int4* pre_list = nullptr; int* pre_n = nullptr; uint8_t* pre_op = nullptr; size_t pre_cap = 0;
  // tuning / test options (coh_set_option; the environment is read once, in coh_init)
  int opt_walk_h = 0;          // 0 = chosen per pass; 1 | 4 | 16 forces the walker's work-item height
  int opt_fused = -1;          // -1 = chosen per pass; 1 fused walker, 0 three-phase frame
  bool aa_general = false;     // every pair through the general (bit-row) antialiasing kernel
  // asynchronous read-back (coh_fb_read_rgba_async): two staging buffers, a copy stream
  cudaStream_t copy_stream = nullptr;
  cudaStream_t aux_stream = nullptr;            // background prefill of a three-phase frame runs beside its scan kernels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint32_t* stage[2] = {nullptr, nullptr}; size_t stage_cap[2] = {0, 0}; bool stage_busy[2] = {false, false};
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  int stage_next = 0;
  int* h_total = nullptr;  // pinned
  // cross-tile carry for fancy fills
  int* queue = nullptr; int n_sms = 0;
  int* carry_done = nullptr; int* carry_cnt = nullptr; int2* carry_ent = nullptr;
  size_t carry_slots = 0; int epoch = 0;
  bool own_stream = true, own_fb = true;
  uint32_t* shared_fb = nullptr;            // coh_fb_alloc_shared: cudaMalloc'ed, exported by CUDA IPC
  std::vector<void*> opened_peers;          // coh_fb_open_peer: mappings to close at shutdown
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
      cudaGetLastError(); /* a refused call must not be blamed on the next launch */          \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)
#define FAIL(msg) do { ctx->err = (msg); return 1; } while (0)
// Device memory comes from the stream-ordered pool (cudaMallocAsync): allocation and release are
// ordered on the context's stream and cost microseconds instead of a device-wide synchronisation.
#define DMALLOC(ptr, bytes) cudaMallocAsync((void**)(ptr), (bytes), ctx->stream)
#define DFREE(ptr) do { if (ptr) { cudaFreeAsync((void*)(ptr), ctx->stream); (ptr) = nullptr; } } while (0)
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

// Launch with (or without) the programmatic-stream-serialization attribute: the kernel may be scheduled before the kernel
// ahead of it in the stream has completed and waits for it in cudaGridDependencySynchronize ().
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(void (*kern)(KArgs...), int grid, int block, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.stream = st;
  cudaLaunchAttribute at[1]; memset(at, 0, sizeof at);
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
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
  if (const char* e = getenv("COH_WALK_H")) coh_set_option(ctx, "walk_h", atoi(e));
  if (const char* e = getenv("COH_FUSED")) coh_set_option(ctx, "fused", atoi(e));
  if (const char* e = getenv("COH_AB")) coh_set_option(ctx, "ab", atoi(e));
  if (const char* e = getenv("COH_AA_GENERAL")) coh_set_option(ctx, "aa_general", atoi(e));
  *out = ctx;
  return 0;
}
int coh_set_option(coh_ctx* ctx, const char* name, int32_t value) {
  const std::string n = name ? name : "";
  if (n == "walk_h") { if (value != 0 && value != 1 && value != 4 && value != 16) FAIL("coh_set_option: walk_h is 0, 1, 4 or 16"); ctx->opt_walk_h = value; }
  else if (n == "fused") { if (value < -1 || value > 1) FAIL("coh_set_option: fused is -1, 0 or 1"); ctx->opt_fused = value; }
  else if (n == "aa_general") ctx->aa_general = value != 0;
  else if (n == "bin_cache") ctx->opt_bin_cache = value != 0;
  else if (n == "pre_min_pairs") ctx->opt_pre_min_pairs = value;
  else if (n == "pre_min_pairs_walk") ctx->opt_pre_min_pairs_walk = value;
  else if (n == "comp_rows") ctx->opt_comp_rows = value != 0;
  else if (n == "ab") ctx->opt_ab = value;
  else if (n == "fork_prefill") ctx->opt_fork_prefill = value != 0;
  else if (n == "pdl") ctx->opt_pdl = value != 0;
  else FAIL("coh_set_option: unknown option '" + n + "'");
  return 0;
}

static void free_binset(coh_ctx* ctx, BinSet& b) {
  DFREE(b.cell_head); DFREE(b.cell_rng); DFREE(b.cell_order); DFREE(b.comp_order); DFREE(b.cell_items); DFREE(b.item_cell); DFREE(b.item_attr); DFREE(b.item_rec); DFREE(b.state);
  b = BinSet();
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
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); }
  coh_cache_clear(ctx);
  for (void* p : ctx->opened_peers) cudaIpcCloseMemHandle(p);
  if (ctx->shared_fb) { if (ctx->fb == ctx->shared_fb) { ctx->fb = nullptr; ctx->own_fb = true; } cudaFree(ctx->shared_fb); ctx->shared_fb = nullptr; }
  DFREE(ctx->d_aa); DFREE(ctx->d_error); cudaFreeHost(ctx->h_error); cudaFreeHost(ctx->h_total);
  if (ctx->own_fb) DFREE(ctx->fb);
  DFREE(ctx->u_out); DFREE(ctx->u_init);
  free_binset(ctx, ctx->bins); DFREE(ctx->queue);
  DFREE(ctx->coarse_items); DFREE(ctx->coarse_counts); DFREE(ctx->coarse_off);
  DFREE(ctx->pre_sc); DFREE(ctx->pre_list); DFREE(ctx->pre_n); DFREE(ctx->pre_op);
  DFREE(ctx->carry_done); DFREE(ctx->carry_cnt); DFREE(ctx->carry_ent);
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
  if ((cudaStream_t)stream == ctx->stream) return 0;   // already running on it
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

#include "host_shapes.inl"
#include "host_polygon.inl"
#include "host_scene.inl"
#include "host_render.inl"
#include "host_cache.inl"

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
  return check_error_flag(ctx, "coh_fb_read_rgba (a frame rendered before this read)");  // synchronises; a frame that overflowed a kernel-side limit must not be handed back as good
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
  bool any = false;
  for (int k = 0; k < 2; k++)
    if (ctx->stage_busy[k]) { CK(cudaEventSynchronize(ctx->ev_done[k])); ctx->stage_busy[k] = false; any = true; }
  return any ? check_error_flag(ctx, "coh_fb_read_wait (a frame rendered before these reads)") : 0;
}
// The result sprite of Render.render_frame over `update` (render.mli:211-217 returns a Sprite.sprite): the
// framebuffer's pixels on the update shape, one RGBA8 word per pixel in canonical span order.
int coh_fb_read_sprite(coh_ctx* ctx, coh_shape_t update, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  if (!ctx->fb) FAIL("coh_fb_read_sprite: no framebuffer");
  if (!update) return 0;   // NullShape -> NullSprite
  DevShape* s = (DevShape*)update;
  if (s->card > cap) FAIL("coh_fb_read_sprite: buffer too small");
  if (s->bx0 < 0 || s->by0 < 0 || s->bx1 >= ctx->fr.W || s->by1 >= ctx->fr.H) FAIL("coh_fb_read_sprite: shape outside the framebuffer");
  int* d_off = nullptr; uint32_t* d_out = nullptr;
  if (shape_pixel_offsets(ctx, s, &d_off)) return 1;
  CK(DMALLOC(&d_out, 4 * (size_t)std::max<long long>(s->card, 1)));
  k_gather_spans<uint32_t><<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, d_off, s->n_rows, 0, ctx->fr.W, ctx->fb + (size_t)s->y0 * ctx->fr.W, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)s->card, cudaMemcpyDeviceToHost, ctx->stream));
  const int rc = check_error_flag(ctx, "coh_fb_read_sprite (a frame rendered before this read)");
  DFREE(d_off); DFREE(d_out);
  if (!rc) *n_out = s->card;
  return rc;
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
  const int rc = check_error_flag(ctx, "coh_fb_read_rgb888 (a frame rendered before this read)");
  DFREE(d);
  return rc;
}
// Wxgui.refresh_window (wxgui.ml:333-366) without the canvas walk: the marshalled "RefreshWindow" message of a dirty
// rectangle, its pixel string filled from the framebuffer by coh_fb_read_rgb888's kernel and copy.
int coh_wire_refresh_window(coh_ctx* ctx, int32_t window, int32_t xmin, int32_t ymin, int32_t xmax, int32_t ymax, uint8_t* out, int64_t cap, int64_t* len) {
  if (!len) FAIL("coh_wire_refresh_window: len is NULL");
  *len = 0;
  uint8_t hdr[64]; int32_t hl = 0;
  const int64_t total = coh_host_wire_refresh_window(window, xmin, ymin, xmax, ymax, hdr, &hl);
  if (total < 0) FAIL("coh_wire_refresh_window: not a rectangle (wxgui.ml:335)");
  if (total == 0) return 0;                       // zero-width or zero-height rectangles just do nothing (wxgui.ml:354-357)
  *len = total;
  if (!out || cap < total) return 0;              // the size is reported; nothing is written
  if (coh_fb_read_rgb888(ctx, xmin, ymin, xmax - xmin + 1, ymax - ymin + 1, out + hl)) { *len = 0; return 1; }
  memcpy(out, hdr, (size_t)hl);
  return 0;
}

}  // extern "C"
#include "host_multi.inl"
