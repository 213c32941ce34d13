// kernels_walk.cuh — part of kernels.cuh (included inside namespace coh, in order): the fused front-to-back walker (WalkParams, aa_tile, walk_cell, k_walk).

// ------------------------------------------------------------------------------------
// The fused walker.
// ------------------------------------------------------------------------------------
constexpr int COH_MAX_PEERS = 7;
struct WalkParams {
  const ObjRec* objs;
  const EdgeRec* edges;
  const int* rowedge_ptr;      // K1 edge binning: per (path object, pixel row) candidate edge lists (CSR)
  const int* rowedge_idx;
  const int2* points;          // brush stamp centres (object frame), list order
  const int2* brush_ranges;    // per (stroke, cell of its box): first / last stamp index reaching the cell
  const uint32_t* conv_bits;   // Convolved objects: shape / minshape bit-rows
  const uint32_t* conv_px;     // Convolved objects: pre-convolved RGBA8 canvases
  const uint8_t* stamps;       // brush alpha stamps
  const int2* cell_rng;        // per cell [start, end) into cell_items
  const int* cls_cells;        // cells by list-length class [BIN_CLASSES][n_cells] (heavy first), with
  const int* cls_cnt;          // ... the number of cells in every class; null: row-major order
  const int* cell_items;
  const int2* cell_head;       // per cell: {colour, 1 | 2 (scene list)} when the cell is one opaque covering primitive, else {0, 0}; may be null
  const AATable* aa;
  Frame fr;
  int cell_row0;               // first cell row covered by cell_rng
  int ux0, uy0, ux1, uy1;      // update box, inclusive
  const uint32_t* u_init;      // optional update set as a bit-frame (fr.H x fr.tiles_x words), else box
  uint32_t* u_out;             // optional: `u` after the scene pass (same layout)
  uint32_t* fb;                // RGBA8 framebuffer, fr.W x fr.H
  int* error_flag;             // set to 1 when an object overflows COH_MAXX crossings
  // Band gather fused into the walk: the same framebuffers on the other GPUs of the box (peer-mapped over
  // NVLink); every final pixel is stored to all of them as it is produced, so the strips arrive while the
  // walk is still running and no collective follows.
  uint32_t* peer_fb[COH_MAX_PEERS];
  int n_peers;
  // Three-phase frames (k_walk<..., PRE = true>): scan conversion and antialiasing were done by
  // k_pre_scan / k_pre_vis / k_pre_aa for every (cell item, row) pair; the walk only composites.
  const int* item_cell;        // cell of every entry of cell_items
  const uint2* pre_sc;         // per pair (item * CELL_H + row of the cell): shape / coverage words
  const uint8_t* pre_op;       // per pair: 32 opacity bytes (valid where the pair's edge mask is set)
  int write_clear;             // write clear pixels of the update too (1) or only touched pixels
  int resume;                  // continue a frame: the root accumulators start from what `fb` already holds
  uint32_t* touched;           // optional bit-frame: receives every pixel some object of the pass was composited at (the
                               // shape of the pass's sprite, which is more than its non-clear pixels; Brush.smear needs it)
  // Cross-tile carry for fancy fills (k_walk<true> only): an AA pixel takes the fill at the first
  // x of its span (polygon.ml:736) and a span may begin in a tile further left.  Every tile
  // publishes, per fancy object whose visible edge run touches its right border, where that run
  // began; the tile to its right looks it up.  With fancy fills cells are taken from the queue
  // in row-major order, so the cell waited on has always been started by a resident warp
  // (decoupled look-back).
  int* queue;                  // work queue head: persistent warps take cells with atomicAdd
  int n_cells;
  int* carry_done;             // per (band row, tile): == epoch when the tile has finished
  int* carry_cnt;              // per (band row, tile): number of published entries
  int2* carry_ent;             // per (band row, tile): CARRY_CAP entries (object index, start x)
  int epoch;
};
constexpr int CARRY_CAP = 8;

// Optional phase timing (tools only; -DCOH_PHASE_PROFILE): cycles per phase summed over warps.
#ifdef COH_PHASE_PROFILE
__device__ unsigned long long g_phase_cycles[16];
__device__ unsigned int g_cell_cycles[1 << 20];
#define PH_DECL long long ph_t = clock64(); long long ph_acc[6] = {0, 0, 0, 0, 0, 0};
#define PH_MARK(i) { long long t_ = clock64(); ph_acc[i] += t_ - ph_t; ph_t = t_; }
#define PH_FLUSH() { if (lane == 0) { for (int i_ = 0; i_ < 6; i_++) atomicAdd(&g_phase_cycles[i_], (unsigned long long)ph_acc[i_]); } }
#else
#define PH_DECL
#define PH_MARK(i)
#define PH_FLUSH()
#endif
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  return v;
}

// AA opacity of the visible edge pixels `edge` (bit b = pixel xx0 + b of row yy, object
// frame) of one polygon.  Lane j scan-converts scaled row 16*yy - 32 + j of the x16 edge
// list (polygon.ml:673-692) into its private 544-bit row in shared memory; then for every
// edge pixel the 32 lanes each weigh their row's 32-column window and the warp reduces.
// Returns the opacity of pixel `lane` (undefined where edge bit is 0).
// General (rare) path of the AA scan: this lane's super-sampled row with the 16-entry lists in local
// memory, written into its 544-bit shared-memory row.  Out of line: it must not bloat the hot path.
__device__ __noinline__ bool aa_rows_general(const EdgeRec* __restrict__ edges, const int* __restrict__ idx, int n_cand,
                                             int winding, int yy, int lane, uint32_t* row, int wlo, int whi) {
  for (int w = 0; w < AA_WORDS; w++) row[w] = 0u;
  SinkMem sink; sink.wx0 = wlo; sink.nwords = AA_WORDS; sink.stride = 1; sink.S = row; sink.C = nullptr;
  ScanState st;
  scan_begin(st, 16 * yy - 32 + lane, true, wlo, whi);
  for (int i = 0; i < n_cand; i++) {
    const EdgeRec e = edges[idx ? idx[i] : i];
    const int x0 = e.x0in * 16, x1 = e.x1in * 16;
    scan_edge(st, x0, x1, e.ymin * 16, e.ymax * 16, e.g, e.dir, edge_side(x0, x1, wlo, whi), sink);
  }
  const bool ok = scan_finish(st, winding, sink);
  __syncwarp();
  return ok;
}
// Staged edge of the AA scan: scaled coordinates plus where it lies relative to the window.
struct StagedEdge { int x0, x1, ymin, ymax; double g; int dir, side; };
constexpr int STAGE_WORDS = sizeof(StagedEdge) / 4;
__device__ __forceinline__ int aa_tile(const EdgeRec* __restrict__ edges, const int* __restrict__ idx, int n_cand,
                                    int winding, int xx0, int yy, uint32_t edge,
                                    uint32_t* aa_bits /*32*AA_WORDS, warp private*/,
                                    StagedEdge* stage /*32, warp private*/,
                                    const int* __restrict__ prefix /*[32][33] shared*/, int volume, int lane,
                                    bool& ok) {
#ifdef COH_PHASE_PROFILE
  long long t0_ = clock64();
#endif
  uint32_t* row = aa_bits + lane * AA_WORDS;
  const int wlo = 16 * xx0 - 32, whi = wlo + 32 * AA_WORDS - 1;
  // Fast path: the crossings of this lane's row stay in registers (a row of a 34-pixel window is
  // touched by one or two edges); if any lane needs more, the whole warp redoes the row with the
  // general lists in local memory.
  constexpr int FAST_X = 3;
  ScanStateT<FAST_X, true> fst;
  SinkRow fsink; fsink.wx0 = wlo; fsink.nwords = AA_WORDS; fsink.saddr = (uint32_t)__cvta_generic_to_shared(row);
#pragma unroll
  for (int w = 0; w < AA_WORDS; w++) row[w] = 0u;
  // Only the super-sampled columns under the edge pixels are ever read back: classify and rank crossings
  // against that narrower window (pixel b reads columns wlo + 16 b .. wlo + 16 b + 31).  Everything to
  // its left only contributes a winding count, everything to its right only "a successor exists".
#ifndef COH_AA_WIDE
  const int nlo = wlo + 16 * (__ffs((int)edge) - 1), nhi = wlo + 16 * (31 - __clz((int)edge)) + 31;
#else
  const int nlo = wlo, nhi = whi;
#endif
  scan_begin(fst, 16 * yy - 32 + lane, true, nlo, nhi);
  // The candidate edges are the same for all 32 super-sampled rows: lane i fetches and scales
  // candidate i once (one parallel round trip to L2 instead of a dependent chain per lane) and
  // classifies it against the window; then every lane walks the staged copies in shared memory.
  for (int base = 0; base < n_cand; base += 32) {
    const int i = base + lane;
    if (i < n_cand) {
      const EdgeRec e = edges[idx ? idx[i] : i];
      StagedEdge se;
      se.x0 = e.x0in * 16; se.x1 = e.x1in * 16; se.ymin = e.ymin * 16; se.ymax = e.ymax * 16;
      se.g = e.g; se.dir = e.dir; se.side = edge_side(se.x0, se.x1, nlo, nhi);
      stage[lane] = se;
    }
    __syncwarp();
#ifdef COH_PHASE_PROFILE
    if (lane == 0) atomicAdd(&g_phase_cycles[12], (unsigned long long)(clock64() - t0_));
#endif
    const int cnt = min(32, n_cand - base);
    for (int k = 0; k < cnt; k++) {
      const StagedEdge se = stage[k];
      scan_edge(fst, se.x0, se.x1, se.ymin, se.ymax, se.g, se.dir, se.side, fsink);
    }
    __syncwarp();
  }
#ifdef COH_PHASE_PROFILE
  if (lane == 0) atomicAdd(&g_phase_cycles[13], (unsigned long long)(clock64() - t0_));
#endif
  bool fast = scan_finish(fst, winding, fsink);
  fast = __all_sync(0xFFFFFFFFu, fast);
#ifdef COH_PHASE_PROFILE
  if (lane == 0) atomicAdd(&g_phase_cycles[14], (unsigned long long)(clock64() - t0_));
#endif
  ok = true;
  __syncwarp();
  if (!fast) ok = aa_rows_general(edges, idx, n_cand, winding, yy, lane, row, wlo, whi);
#ifdef COH_PHASE_PROFILE
  long long t1_ = clock64();
  if (lane == 0) { atomicAdd(&g_phase_cycles[6], (unsigned long long)(t1_ - t0_)); atomicAdd(&g_phase_cycles[8], 1ull); atomicAdd(&g_phase_cycles[9], (unsigned long long)n_cand); atomicAdd(&g_phase_cycles[10], (unsigned long long)__popc(edge)); atomicAdd(&g_phase_cycles[11], fast ? 0ull : 1ull); }
#endif
  int mytot = 0;
  const int* prow = prefix + lane * 33;
  uint32_t e = edge;
  // Four edge pixels per iteration: a pixel's table sum is at most 42120 < 2^16, so two pixels
  // share one 32-bit register through the butterfly reduction, and two such registers are
  // reduced side by side (independent shuffles pipeline).
  while (e) {
    int b[4];
    int part[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      b[k] = e ? (__ffs((int)e) - 1) : -1;
      e &= e - 1;   // 0 & anything stays 0
      part[k] = 0;
      if (b[k] >= 0) {
        const uint32_t lo = row[b[k] >> 1], hi = row[(b[k] >> 1) + 1];
        const uint32_t m = (b[k] & 1) ? ((lo >> 16) | (hi << 16)) : lo;
        part[k] = aa_row_sum(prow, m);
      }
    }
    int p01 = part[0] | (part[1] << 16), p23 = part[2] | (part[3] << 16);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      p01 += __shfl_xor_sync(0xFFFFFFFFu, p01, d);
      p23 += __shfl_xor_sync(0xFFFFFFFFu, p23, d);
    }
    mytot = (lane == b[0]) ? (p01 & 0xFFFF) : mytot;
    mytot = (lane == b[1]) ? ((p01 >> 16) & 0xFFFF) : mytot;
    mytot = (lane == b[2]) ? (p23 & 0xFFFF) : mytot;
    mytot = (lane == b[3]) ? ((p23 >> 16) & 0xFFFF) : mytot;
  }
  __syncwarp();
#ifdef COH_PHASE_PROFILE
  if (lane == 0) { atomicAdd(&g_phase_cycles[7], (unsigned long long)(clock64() - t1_)); }
#endif
  return aa_opacity(mytot, volume);  // one division per lane, after the loop
}

// Shape (S) and coverage (C) words of one pixel row of one path object inside a 32-pixel window.
// Crossings stay in registers (3 per list); a row with more takes the general lists.
__device__ __noinline__ uint2 scan_row_word(const EdgeRec* __restrict__ edges, const int* __restrict__ idx, int n_cand,
                                            int yy, int winding, int xx0, bool& ok) {
  ScanStateT<3, true> st;
  Sink32 sink; sink.wx0 = xx0; sink.S = 0u; sink.C = 0u;
  scan_begin(st, yy, false, xx0, xx0 + 31);
  for (int i = 0; i < n_cand; i++) {
    const EdgeRec e = edges[idx[i]];
    scan_edge(st, e.x0in, e.x1in, e.ymin, e.ymax, e.g, e.dir, edge_side(e.x0in, e.x1in, xx0, xx0 + 31), sink);
  }
  if (!scan_finish(st, winding, sink)) {
    sink.S = 0u; sink.C = 0u;
    if (!scan_row(edges, idx, n_cand, 1, yy, winding, false, xx0, xx0 + 31, sink)) ok = false;
  }
  return make_uint2(sink.S, sink.C);
}
// 32 bits of a bit-row starting at an arbitrary bit offset (zeros outside the row)
__device__ __forceinline__ uint32_t conv_load_bits32(const uint32_t* __restrict__ row, int nw, int bitoff) {
  const int qw = bitoff >> 5, qb = bitoff & 31;
  const uint32_t lo = (qw >= 0 && qw < nw) ? row[qw] : 0u;
  const uint32_t hi = (qw + 1 >= 0 && qw + 1 < nw) ? row[qw + 1] : 0u;
  return qb ? ((lo >> qb) | (hi << (32 - qb))) : lo;
}
// out-of-line copy for the rarer object kinds (keeps the polygon walker's code small)
__device__ __noinline__ int aa_tile_nl(const EdgeRec* __restrict__ edges, const int* __restrict__ idx, int n_cand, int winding,
                                       int xx0, int yy, uint32_t edge, uint32_t* aa_bits, StagedEdge* stage,
                                       const int* __restrict__ prefix, int volume, int lane, bool& ok) {
  return aa_tile(edges, idx, n_cand, winding, xx0, yy, edge, aa_bits, stage, prefix, volume, lane, ok);
}
constexpr int WALK_WARPS = 8;            // warps (= cells) per CTA
#ifndef WALK_MIN_CTAS
#define WALK_MIN_CTAS 3
#endif
// A cell list (32 px x CELL_H rows) is shared by CELL_H / WALK_H walker work items of WALK_H rows
// each: the heaviest work item bounds the kernel's critical path, so for scenes with short lists and
// heavy antialiasing (the lion) the unit of work is 4 rows; scenes with very long lists (10^5
// objects) amortise the list walk over all 16 rows.  WALK_H is a template parameter of the walker.
// One warp owns one cell: TILE_W = 32 pixel columns (lane = column when compositing) by
// WALK_H rows.  Scan conversion runs lane-parallel over (candidate object, row) pairs; the
// front-to-back composite then visits, object by object, only the rows where the object
// still has pixels inside the covered-so-far complement `u` (one 32-bit word per row, held
// by lane r and its NC-1 mirror lanes).
// CARRY: the scene has fancy (gradient / radial) fills -> fill evaluation and the cross-tile carry
// are compiled in.  EXTRAS: 0 = polygons and primitives only (the small kernel), 1 = + brush strokes and
// Convolved objects, 2 = + CPG objects and continuing a frame (filter passes).
template <bool CARRY, int EXTRAS, int WALK_H, bool PRE>
__device__ __forceinline__ void walk_cell(const WalkParams& P, const int tile, const int by, const int sub, const int lane,
                                          uint32_t* __restrict__ aa_bits, StagedEdge* __restrict__ stage,
                                          uint32_t (*__restrict__ acc_rows)[32],
                                          const int* __restrict__ s_prefix, const int volume) {
  constexpr bool BRUSH = EXTRAS >= 1;      // brush strokes, Convolved objects
  constexpr bool CPGX = EXTRAS >= 2;       // CPG objects, continuing a frame (filters)
  constexpr int NC = 32 / WALK_H;          // candidate objects scan-converted per pass
  constexpr unsigned ROWMASK = (WALK_H >= 32) ? 0xFFFFFFFFu : ((1u << WALK_H) - 1u);
  const int tx0 = tile * TILE_W;
  const int y0 = (P.cell_row0 + by) * CELL_H + sub * WALK_H;
  const int r_lane = lane % WALK_H, c_lane = lane / WALK_H;
  const int my_y = y0 + r_lane;                      // the row whose `u` this lane mirrors
  const bool row_in_band = my_y >= P.fr.band_y0 && my_y < P.fr.band_y1;
  const size_t my_slot = (size_t)(my_y - P.fr.band_y0) * P.fr.tiles_x + tile;  // carry slot of (row, tile)
  int n_carry = 0;                                   // published carry entries of my row (mirrored)

  const int cell = by * P.fr.cntx + tile - P.fr.ctx0;
  const int2 cell_rg = P.cell_rng[cell];
  const int it0 = cell_rg.x, it1 = cell_rg.y;
  const int2 head = (P.cell_head && !(CPGX && P.resume)) ? P.cell_head[cell] : make_int2(0, 0);
  // initial covered-so-far complement `u` of my row's word
  uint32_t u = 0u;
  if (row_in_band) {
    if (P.u_init) u = P.u_init[(size_t)my_y * P.fr.tiles_x + tile];
    else u = (my_y >= P.uy0 && my_y <= P.uy1) ? interval_mask32(tx0, P.ux0, P.ux1) : 0u;
    if (tx0 + 31 >= P.fr.W) u &= interval_mask32(tx0, 0, P.fr.W - 1);
  }
  const uint32_t u_update = u;
  if (P.u_out && row_in_band && c_lane == 0) P.u_out[(size_t)my_y * P.fr.tiles_x + tile] = u;  // nothing covered yet
  auto publish_done = [&]() {
    if (CARRY && row_in_band && c_lane == 0) {
      P.carry_cnt[my_slot] = n_carry < CARRY_CAP ? n_carry : CARRY_CAP;
      __threadfence();
      atomicExch(P.carry_done + my_slot, P.epoch);
    }
  };
  if (__ballot_sync(0xFFFFFFFFu, u != 0u) == 0u) { publish_done(); return; }

  // Fast path: the only object reaching this cell is an opaque primitive that covers all of it
  // (typically the background rectangle; flagged by the binning kernel): the rows are just that colour.
  if (head.y & 1) {
    const uint32_t c0 = (uint32_t)head.x;
    if (CPGX && P.touched && row_in_band && c_lane == 0 && u) P.touched[(size_t)my_y * P.fr.tiles_x + tile] |= u;   // (filter scenes only: they walk with EXTRAS = 2)
    if ((head.y & 2) && P.u_out && row_in_band && c_lane == 0) P.u_out[(size_t)my_y * P.fr.tiles_x + tile] = 0u;
    publish_done();
#pragma unroll 1
    for (int r = 0; r < WALK_H; r++) {
      const uint32_t uu = __shfl_sync(0xFFFFFFFFu, u, r);
      if ((uu >> lane) & 1u) {
        const size_t at = (size_t)(y0 + r) * P.fr.W + tx0 + lane;
        P.fb[at] = c0;
        for (int k = 0; k < P.n_peers; k++) P.peer_fb[k][at] = c0;
      }
    }
    return;
  }

#pragma unroll
  for (int r = 0; r < WALK_H; r++) acc_rows[r][lane] = 0u;   // accumulator of the current nesting level
  __syncwarp();
  uint32_t touched_w = 0u;             // pixels of my row some object was composited at
  int depth = 0;                       // open groups
  int hit_level = -1;                  // outermost open group that dissolves its sprite (PreTrans), or -1
  int open_grp[MAX_DEPTH];
  uint32_t stk_u[MAX_DEPTH];           // parents' `u` of my row
  uint32_t stk_acc[MAX_DEPTH][WALK_H]; // parents' accumulators of my column (local memory; touched on push/pop only)
  bool bad = false;

  PH_DECL

  auto pop_group = [&]() {
    // close the innermost group: its accumulated sprite goes under the parent accumulator
    // (render.ml:1294 caf over opaque a s; 1295-1298 PreTrans), newly opaque pixels leave the
    // parent's u (render.ml:1308).
    const int g = open_grp[depth - 1];
    const int pt = P.objs[g].pretrans;
    const int gflags = P.objs[g].flags;
    const uint32_t pu = stk_u[depth - 1];
#pragma unroll 1
    for (int r = 0; r < WALK_H; r++) {
      uint32_t sp = acc_rows[r][lane];
      const uint32_t pa = stk_acc[depth - 1][r];
      if (pt >= 0) sp = px_dissolve(sp, pt);
      const uint32_t res = px_over(pa, sp);
      const uint32_t opq = __ballot_sync(0xFFFFFFFFu, (res >> 24) == 255u);
      acc_rows[r][lane] = res;
      if (r_lane == r) u = pu & ~opq;
    }
    depth--;
    if (hit_level >= depth) hit_level = -1;
    if ((gflags & OF_ROOT_SCENE) && P.u_out && row_in_band && c_lane == 0) P.u_out[(size_t)my_y * P.fr.tiles_x + tile] = u;
  };
  auto push_group = [&](int g) {
    // Continuing a frame (filter passes, render.ml:1080-1131): the scene list's accumulator carries on from the
    // framebuffer (render_scene's `a`); the background list is composited under it (render.ml:1363-1365).
    const int rflags = (CPGX && P.resume && depth == 0) ? P.objs[g].flags : 0;
#pragma unroll 1
    for (int r = 0; r < WALK_H; r++) {
      uint32_t below = acc_rows[r][lane], fresh = 0u;
      if (CPGX && rflags) {
        const int py = y0 + r, px = tx0 + lane;
        const uint32_t have = (py < P.fr.H && px < P.fr.W) ? P.fb[(size_t)py * P.fr.W + px] : 0u;
        if (rflags & OF_ROOT_SCENE) fresh = have; else below = have;
      }
      stk_acc[depth][r] = below; acc_rows[r][lane] = fresh;
    }
    stk_u[depth] = u;
    open_grp[depth] = g;
    // A group composited with PreTrans (v < 1) gives pixels back to its parent's `u` when it
    // closes (an opaque member pixel is no longer opaque once dissolved), so while such a group
    // is open, candidates are pre-selected against the u saved outside the outermost one.
    const int pt = P.objs[g].pretrans;
    if (hit_level < 0 && pt >= 0 && pt < 255) hit_level = depth;
    depth++;
  };

  // the loop runs one extra, empty pass whose only effect is to close every open group
  // (keeps a single inlined copy of the group push/pop code: the kernel must fit the I-cache)
  for (int base = it0;; base += NC) {
    const uint32_t u_hit = hit_level >= 0 ? stk_u[hit_level] : u;  // superset of every later u of my row
    // nothing of these rows is uncovered any more: the rest of the list cannot show
    // (render.ml:1321-1322: render_scene stops when u is null)
    if (it1 - base > 2 * NC && __ballot_sync(0xFFFFFFFFu, u_hit != 0u) == 0u) base = it1;
    const bool closing = base >= it1;
    PH_MARK(0)  // other / loop overhead
    // ---- lane-parallel scan conversion: lane (c, r) evaluates row y0+r of candidate c ----
    const int ci = base + c_lane;
    const int idx = ci < it1 ? P.cell_items[ci] : -1;
    uint32_t S = 0u, C = 0u;
    uint32_t gSA = 0u, gMA = 0u, gSB = 0u, gMB = 0u;   // CPG operands: shape / minshape words of a and b
    if (PRE) {
      if (idx >= 0 && u_hit != 0u) {   // scan-converted by k_pre_scan: no object, row list or edge is touched here
        const uint2 sc = P.pre_sc[(size_t)ci * CELL_H + sub * WALK_H + r_lane];
        S = sc.x; C = sc.y;
      }
    } else if (idx >= 0 && u_hit != 0u) {
      const ObjRec& o = P.objs[idx];
      if (!(o.by0 > my_y || o.by1 < my_y || o.bx0 > tx0 + 31 || o.bx1 < tx0)) {
        const int yy = my_y - o.dy, xx0 = tx0 - o.dx;
        if (o.kind == K_PRIM) {
          if (yy >= o.prim[1] && yy <= o.prim[3]) S = interval_mask32(xx0, o.prim[0], o.prim[2]);
        } else if (o.kind == K_PATH) {
          if (yy >= o.ry0 && yy <= o.ry1) {
            const int slot = o.row_base + yy - o.ry0;
            const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
            bool ok = true;
            const uint2 sc = scan_row_word(P.edges, P.rowedge_idx + a, b - a, yy, o.winding, xx0, ok);
            if (!ok) bad = true;
            S = sc.x; C = sc.y;
          }
        } else if (CPGX && o.kind == K_CPG) {
          // CPG (op, a, b): shape / minshape are set expressions of the operands' (render.ml:522-528)
          bool ok = true;
          if (yy >= o.ry0 && yy <= o.ry1) {
            const int slot = o.row_base + yy - o.ry0;
            const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
            const uint2 sc = scan_row_word(P.edges, P.rowedge_idx + a, b - a, yy, o.winding, xx0, ok);
            gSA = sc.x; gMA = sc.x & ~sc.y;
          }
          if (yy >= o.b_ry0 && yy <= o.b_ry1) {
            const int slot = o.b_row_base + yy - o.b_ry0;
            const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
            const uint2 sc = scan_row_word(P.edges, P.rowedge_idx + a, b - a, yy, o.b_opw >> 8, xx0, ok);
            gSB = sc.x; gMB = sc.x & ~sc.y;
          }
          if (!ok) bad = true;
          uint32_t M;
          cpg_words(o.b_opw & 255, gSA, gMA, gSB, gMB, S, M);
          C = S & ~M;
        } else if (BRUSH && o.kind == K_CONV) {
          // Convolved (k, g): shape = bloat r r (shape g), minshape = erode r r (minshape g) (render.ml:536-555),
          // both precomputed as bit-rows; C is chosen so that S & ~C is the minshape word
          if (yy >= o.cv_y0 && yy < o.cv_y0 + o.cv_h) {
            const uint32_t* rowS = P.conv_bits + o.cv_bits + (size_t)(yy - o.cv_y0) * o.cv_nw;
            S = conv_load_bits32(rowS, o.cv_nw, xx0 - o.cv_x0);
            C = S & ~conv_load_bits32(rowS + (size_t)o.cv_h * o.cv_nw, o.cv_nw, xx0 - o.cv_x0);
          }
        } else if (BRUSH && o.kind == K_BRUSH) {
          // shape = dilation of the stamp centres by the brush box (brush.ml:143-168); minshape null
          const int br = o.brush_r;
          const int2 rg = brush_range(P.brush_ranges, o, xx0, yy);
          for (int k = rg.x; k <= rg.y; k++) {
            int2 p = P.points[o.first + k];
            if (p.y - br <= yy && yy <= p.y + br) S |= interval_mask32(xx0, p.x - br, p.x + br);
          }
          C = S;
        }
      }
    }
    PH_MARK(1)  // scan
    const unsigned hits = __ballot_sync(0xFFFFFFFFu, (S & u_hit) != 0u);
    if (hits == 0u && !closing) continue;
    // ---- sequential front-to-back composite of the candidates that still show ----
    for (int cc = 0; cc < NC; cc++) {
      unsigned rows = (hits >> (cc * WALK_H)) & ROWMASK;
      if (rows == 0u && !(closing && cc == 0)) continue;
      const int ik = closing ? 0 : __shfl_sync(0xFFFFFFFFu, idx, cc * WALK_H);
      const ObjRec& o = P.objs[ik];
      PH_MARK(0)
      // group transitions: close groups that do not enclose this object, open the ones that do
      const int odepth = closing ? 0 : o.depth;
      int common = 0;
      while (common < depth && common < odepth && open_grp[common] == o.anc[common]) common++;
      while (depth > common) pop_group();
      if (closing) break;
      while (depth < odepth) push_group(o.anc[depth]);
      const int okind = o.kind, fkind = o.fill.kind, pretrans = o.pretrans, odx = o.dx, ody = o.dy;
      const uint32_t c0 = o.fill.c0;
      PH_MARK(2)  // transitions
      while (rows) {
        const int r = __ffs((int)rows) - 1;
        rows &= rows - 1;
        const uint32_t Sk = __shfl_sync(0xFFFFFFFFu, S, cc * WALK_H + r);
        const uint32_t Ck = __shfl_sync(0xFFFFFFFFu, C, cc * WALK_H + r);
        const uint32_t ur = __shfl_sync(0xFFFFFFFFu, u, r);
        const uint32_t vis = Sk & ur;
        if (vis == 0u) continue;
        if (CPGX && r_lane == r) touched_w |= vis;
        const uint32_t M = Sk & ~Ck;          // minshape word (polygon.ml:526)
        const uint32_t edge = vis & ~M;       // shptorender ∩ maxshape (render.ml:1201-1204)
        const int yy = y0 + r - ody, xx0 = tx0 - odx;
        int opacity = 255;
        PH_MARK(3)  // row setup
        if (edge) {
          if (PRE) {
            if (okind == K_PATH) opacity = P.pre_op[((size_t)(base + cc) * CELL_H + sub * WALK_H + r) * 32 + lane];
          } else if (okind == K_PATH) {
            bool ok;
            const int slot = o.row_base + yy - o.ry0;
            const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
            opacity = aa_tile(P.edges, P.rowedge_idx + a, b - a, o.aa_winding, xx0, yy, edge, aa_bits, stage, s_prefix, volume, lane, ok);
            if (!ok) bad = true;
          } else if (CPGX && okind == K_CPG) {
            // sprite_of_cpg (render.ml:867-981): both operands become antialiased alpha mattes (0 outside
            // their shape) and are combined per pixel
            const int src = cc * WALK_H + r;
            const uint32_t SA = __shfl_sync(0xFFFFFFFFu, gSA, src), MA = __shfl_sync(0xFFFFFFFFu, gMA, src);
            const uint32_t SB = __shfl_sync(0xFFFFFFFFu, gSB, src), MB = __shfl_sync(0xFFFFFFFFu, gMB, src);
            const uint32_t XA = edge & SA & ~MA, XB = edge & SB & ~MB;
            int a = 0, b = 0;
            bool ok = true;
            if (XA) {
              const int slot = o.row_base + yy - o.ry0;
              const int ea = P.rowedge_ptr[slot], eb = P.rowedge_ptr[slot + 1];
              a = aa_tile_nl(P.edges, P.rowedge_idx + ea, eb - ea, o.winding, xx0, yy, XA, aa_bits, stage, s_prefix, volume, lane, ok);
              if (!ok) bad = true;
            }
            if (XB) {
              const int slot = o.b_row_base + yy - o.b_ry0;
              const int ea = P.rowedge_ptr[slot], eb = P.rowedge_ptr[slot + 1];
              b = aa_tile_nl(P.edges, P.rowedge_idx + ea, eb - ea, o.b_opw >> 8, xx0, yy, XB, aa_bits, stage, s_prefix, volume, lane, ok);
              if (!ok) bad = true;
            }
            // Inside an operand's minshape the reference never consults that operand's matte (regions
            // min/max and max/min take the other operand's alpha, or its inverse): 255 stands for it in
            // cpg_alpha.  (Its sampled value could be below 255: the minshape's row band misses the top
            // quarter of the AA window.)
            a = ((MA >> lane) & 1u) ? 255 : (((XA >> lane) & 1u) ? a : 0);
            b = ((MB >> lane) & 1u) ? 255 : (((XB >> lane) & 1u) ? b : 0);
            opacity = cpg_alpha(o.b_opw & 255, a, b);
          } else if (BRUSH && okind == K_BRUSH) {
            // ordered alpha_over of every stamp covering this pixel (brush.ml:207-212)
            const int br = o.brush_r, w = 2 * br + 1;
            const int px = xx0 + lane;
            uint32_t al = 0u;
            if ((edge >> lane) & 1u) {
              const int2 rg = brush_range(P.brush_ranges, o, xx0, yy);
              for (int q = rg.x; q <= rg.y; q++) {
                int2 p = P.points[o.first + q];
                int ddx = px - p.x, ddy = yy - p.y;
                if (ddx >= -br && ddx <= br && ddy >= -br && ddy <= br)
                  al = alpha_over(al, P.stamps[o.stamp_off + (ddy + br) * w + (ddx + br)]);
              }
            }
            opacity = (int)al;
          }
        }
        PH_MARK(4)  // AA
        int lead_start = xx0;  // object-frame x where the run containing bit 0 begins
        if (CARRY && edge && okind == K_PATH && fkind != 0) {
          const size_t slot = (size_t)(y0 + r - P.fr.band_y0) * P.fr.tiles_x + tile;
          if ((edge & 1u) && tile > P.fr.ctx0) {   // (nothing is visible left of the pass's first column: u is empty there)
            const volatile int* done = P.carry_done + slot - 1;
            while (*done != P.epoch) __nanosleep(32);
            __threadfence();
            const int cnt = P.carry_cnt[slot - 1];
            for (int q = 0; q < cnt; q++) {
              int2 e = P.carry_ent[(slot - 1) * CARRY_CAP + q];
              if (e.x == ik) lead_start = e.y - odx;
            }
          }
          if (edge >> 31) {
            const uint32_t nz = ~edge;
            const int tstart = nz ? (xx0 + 32 - __clz((int)nz)) : lead_start;
            const int nc = __shfl_sync(0xFFFFFFFFu, n_carry, r);
            if (lane == 0 && nc < CARRY_CAP) P.carry_ent[slot * CARRY_CAP + nc] = make_int2(ik, tstart + odx);
            if (r_lane == r) n_carry++;
            if (nc + 1 > CARRY_CAP) bad = true;
          }
        }
        const bool mine = (vis >> lane) & 1u;
        uint32_t acc = acc_rows[r][lane];
        if (mine) {
          const bool is_edge = (edge >> lane) & 1u;
          uint32_t col;
          if (BRUSH && okind == K_CONV && is_edge)  // the convolved sprite, cropped to the visible max-shape (render.ml:1052)
            col = P.conv_px[(size_t)o.cv_px + (size_t)(yy - o.cv_y0) * (o.cv_nw * 32) + (xx0 + lane - o.cv_x0)];
          else if (!CARRY || okind == K_PRIM || fkind == 0) col = c0;
          else if (!is_edge || okind == K_BRUSH || okind == K_CPG) col = fill_lookup(o.fill, xx0 + lane, yy);  // per-pixel fill (brush.ml / render.ml:975 map_coords)
          else {
            // polygon.ml:736 quirk: AA pixels take the fill at the first x of their span (the run
            // of `edge` bits); a run that reaches bit 0 may have begun in a tile further left.
            uint32_t below = ~edge & ((1u << lane) - 1u);
            int start = below ? (xx0 + 32 - __clz((int)below)) : lead_start;
            col = fill_lookup(o.fill, start, yy);
          }
          if (is_edge && !(BRUSH && okind == K_CONV)) col = px_dissolve(col, opacity);
          if (pretrans >= 0) col = px_dissolve(col, pretrans);
          acc = px_over(acc, col);
          acc_rows[r][lane] = acc;
        }
        const uint32_t opq = __ballot_sync(0xFFFFFFFFu, mine && (acc >> 24) == 255u);
        PH_MARK(5)  // composite
        if (r_lane == r) u &= ~opq;  // u' = u --- f  (render.ml:1308)
      }
    }
    if (closing) break;
  }
  PH_MARK(2)
  publish_done();
  if (bad) *P.error_flag = 1;
  if (CPGX && P.touched && row_in_band && c_lane == 0 && touched_w) P.touched[(size_t)my_y * P.fr.tiles_x + tile] |= touched_w;
#pragma unroll 1
  for (int r = 0; r < WALK_H; r++) {
    const uint32_t uu = __shfl_sync(0xFFFFFFFFu, u_update, r);
    if ((uu >> lane) & 1u) {
      const uint32_t acc = acc_rows[r][lane];
      if (P.write_clear || acc != 0u) {
        const size_t at = (size_t)(y0 + r) * P.fr.W + tx0 + lane;
        P.fb[at] = acc;
        for (int k = 0; k < P.n_peers; k++) P.peer_fb[k][at] = acc;
      }
    }
  }
  PH_MARK(0)
  PH_FLUSH()
}

// Persistent launch: every warp keeps taking cells from the queue (heavy cells first) until it
// is empty; the grid is sized to fill the GPU exactly once (WALK_MIN_CTAS CTAs per SM).
template <bool CARRY, int EXTRAS, int WALK_H, bool PRE = false>
__global__ void __launch_bounds__(WALK_WARPS * 32, WALK_MIN_CTAS) k_walk(WalkParams P) {
  constexpr int WALK_SUB = CELL_H / WALK_H;
  __shared__ __align__(16) int s_prefix[32 * 33];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_aa[WALK_WARPS][32 * AA_WORDS];
  __shared__ uint32_t s_acc[WALK_WARPS][WALK_H][32];
  __shared__ StagedEdge s_stage[WALK_WARPS][32];
  stage_aa_table(s_prefix, &s_bar, P.aa);
  const int volume = P.aa->volume;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // Work items come off one atomic counter.  The heavy cells at the head of the order are taken one item at a
  // time (balance); the long tail of cells with at most one object (mostly background) is taken in batches,
  // or every warp of the GPU would queue up on the same counter for ~100 instructions of work per item.
  __shared__ int s_cls[BIN_CLASSES + 1];   // first position of every length class in the heavy-first order
  if (P.cls_cnt && threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < BIN_CLASSES; c++) { s_cls[c] = acc; acc += P.cls_cnt[c]; }
    s_cls[BIN_CLASSES] = acc;
  }
  __syncthreads();
  const int n_items = (P.cls_cnt ? s_cls[BIN_CLASSES] : P.n_cells) * WALK_SUB;   // cells finished by the binning kernel are in no class
  const int heavy_items = P.cls_cnt ? s_cls[BIN_HEAVY_CLASSES] * WALK_SUB : n_items;  // cells with >= 2 objects
  int q_next = 0, q_end = 0, cur_cls = 0;
  for (;;) {
    if (q_next == q_end) {
      const int batch = q_end >= heavy_items ? 8 : 1;
      if (lane == 0) q_next = atomicAdd(P.queue, batch);
      q_next = __shfl_sync(0xFFFFFFFFu, q_next, 0);
      q_end = q_next + batch;
    }
    const int q = q_next++;
    if (q >= n_items) break;
    int cell = q / WALK_SUB;
    if (P.cls_cnt) {
      while (cur_cls < BIN_CLASSES - 1 && cell >= s_cls[cur_cls + 1]) cur_cls++;   // a warp's queue positions only grow
      cell = P.cls_cells[(size_t)cur_cls * P.n_cells + cell - s_cls[cur_cls]];
    }
    const int sub = q % WALK_SUB;
#ifdef COH_PHASE_PROFILE
    long long tc0_ = clock64();
#endif
    walk_cell<CARRY, EXTRAS, WALK_H, PRE>(P, P.fr.ctx0 + cell % P.fr.cntx, cell / P.fr.cntx, sub, lane, s_aa[wid], s_stage[wid], s_acc[wid], s_prefix, volume);
    __syncwarp();
#ifdef COH_PHASE_PROFILE
    if (lane == 0 && cell < (1 << 20)) g_cell_cycles[cell] = (unsigned int)(clock64() - tc0_);
#endif
  }
}
