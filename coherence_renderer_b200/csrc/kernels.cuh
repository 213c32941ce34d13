// kernels.cuh — hand-written sm_100a kernels of the raster hot path.
//
//   k_prep_edges      int32 edges -> EdgeRec (x0in/x1in/ymin/ymax, FP64 gradient, direction)
//   k_rowedges        K1, once per scene: per (path object, pixel row) candidate edge lists (CSR)
//   k_brush_cells     K1, once per scene: per (stroke, cell of its box) range of stamps reaching the cell
//   k_bin1            K1, every frame, <= 1024 leaves: per-cell (32 px x 16 rows) front-to-back object lists in
//                     one pass (ballot compaction, pool cursor), list-length classes for the heavy-first order
//   k_bin_obj / k_bin_sort / k_bin2   K1, every frame, large scenes: two-level binning
//   k_walk            K2+K4+K5 fused: one warp per work item (32 px x 1/4/16 rows of a cell) walks the cell
//                     list front to back; scan-converts each candidate object's row into 32-bit
//                     shape/coverage words, prunes with the covered-so-far word `u`, evaluates the
//                     correlated-matte AA only for still-visible edge pixels, composites with
//                     8-bit premultiplied `over`, and subtracts newly opaque pixels from `u`.
//                     The row of RGBA8 is written once, 128 B per warp, fully coalesced (and mirrored to
//                     the peer framebuffers of the other GPUs when bands are gathered).
//   k_pre_scan / k_pre_vis / k_pre_aa + k_walk<PRE>   three-phase frames for plain polygon scenes
//   k_scan_rows       K2 stand-alone: shape + coverage bit-rows of one edge list (export path)
//   k_aa_rows         K4 stand-alone: AA opacity bytes for the pixels of a bit-frame
//   bit-frame kernels K3: span sets <-> bit-frames, AND/OR/ANDNOT, dilation, run extraction
//   k_conv_pass, k_monochrome, k_filter_matte, k_filter_blend   convolve.ml / filters.ml passes
//
// Design notes are in DESIGN.md.  No tensor cores: this is integer/bit work with a few FP64
// crossings per (edge,row); the relevant roofline is HBM.
#pragma once
#include <cuda_runtime.h>
#include "device_types.cuh"

namespace coh {

constexpr int ORDER_BINS = 256;  // ints of binning state per pass (pool cursor, class counts, work-queue head behind them)

// ------------------------------------------------------------------------------------
// The AA prefix table (4224 bytes, polygon.ml:616-671) into shared memory by one TMA bulk copy (cp.async.bulk,
// completion counted on an mbarrier): one elected thread issues it, the copy engine moves the bytes while the block's
// threads set themselves up, everybody waits on the barrier's phase.  `table` and `s_prefix` are 16-byte aligned and
// the size is a multiple of 16 (the copy unit).  Call from all threads of the block, once, before the table is read.
__device__ __forceinline__ void stage_aa_table(int* s_prefix, unsigned long long* s_bar, const AATable* __restrict__ table) {
  constexpr unsigned BYTES = 32 * 33 * sizeof(int);
  static_assert(BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  const unsigned bar = (unsigned)__cvta_generic_to_shared(s_bar), dst = (unsigned)__cvta_generic_to_shared(s_prefix);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(&table->prefix[0][0]), "r"(BYTES), "r"(bar) : "memory");
  }
  __syncthreads();   // the barrier is initialised (and armed) before anybody polls it
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar) : "memory");
  }
}

__global__ void k_prep_edges(const int4* __restrict__ in, EdgeRec* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 e = in[i];  // x0, y0, x1, y1
  out[i] = make_edge(e.x, e.y, e.z, e.w);
}

// ------------------------------------------------------------------------------------
// K1 edge binning (the reference's active-edge bookkeeping, polygon.ml:541-547): for every
// path object and every pixel row it can touch, the list of candidate edges.  A row's
// candidates are the edges whose y range meets the row's extended band [32y-67, 32y+16]:
// the shape band [32y-47, 32y+16] (polygon.ml:539-540) united with the bands of the 32
// super-sampled rows 16y-32 .. 16y-1 of its AA window (polygon.ml:694-700).
// Built once per scene: count (atomics) -> exclusive scan -> fill (atomic cursors).  The
// order inside a list is irrelevant: ties between crossings cannot change the pixel set.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int floordiv32(int a) { return a >> 5; }               // floor(a / 32)
__device__ __forceinline__ int ceildiv32(int a) { return (a + 31) >> 5; }         // ceil(a / 32)
// One warp per edge, lanes over the rows it reaches (an edge of a page-sized rectangle reaches thousands).
template <bool FILL>
__global__ void k_rowedges(const EdgeRec* __restrict__ edges, const int* __restrict__ edge_obj, int n_edges,
                           const ObjRec* __restrict__ objs, int* __restrict__ counts, const int* __restrict__ ptr,
                           int* __restrict__ idx) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= n_edges) return;
  int oi = edge_obj[e];
  if (oi < 0) return;
  const EdgeRec ed = edges[e];
  int row_base = objs[oi].row_base, ry0 = objs[oi].ry0;
  if (objs[oi].kind == K_CPG && e >= objs[oi].b_first) { row_base = objs[oi].b_row_base; ry0 = objs[oi].b_ry0; }  // operand b
  int ylo = ceildiv32(ed.ymin - 16), yhi = floordiv32(ed.ymax + 67);
  for (int y = ylo + lane; y <= yhi; y += 32) {
    int slot = row_base + y - ry0;
    int k = atomicAdd(&counts[slot], 1);
    if (FILL) idx[ptr[slot] + k] = e;
  }
}

// Owner arrays from ranges: one warp per (first, count, record) range.
__global__ void k_fill_owner(const int4* __restrict__ ranges, int n_ranges, int* __restrict__ owner) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n_ranges) return;
  const int4 rg = ranges[r];
  for (int k = lane; k < rg.y; k += 32) owner[rg.x + k] = rg.z;
}
__global__ void k_fill_int2(int2* __restrict__ dst, int n, int2 v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
// K1 for brush strokes: for every (stroke, 32 x CELL_H pixel cell of its box, object frame) the range
// [imin, imax] of stamp indices (list order) whose footprint reaches the cell.  Stamps are sampled along the
// path, so the stamps that reach a cell are (nearly) consecutive; walking the range in order keeps the
// reference's stamping order (brush.ml:207-212) and makes the per-pixel cost independent of the stroke length.
__device__ __forceinline__ int floordiv_pos(int a, int d) { return a >= 0 ? a / d : -((-a + d - 1) / d); }
__global__ void k_brush_cells(const int2* __restrict__ points, const int* __restrict__ point_obj, int n_points,
                              const ObjRec* __restrict__ objs, int2* __restrict__ ranges) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const int oi = point_obj[i];
  if (oi < 0) return;
  const ObjRec& o = objs[oi];
  const int k = i - o.first, r = o.brush_r;
  const int2 p = points[i];
  const int cx0 = floordiv_pos(p.x - r, 32) - o.bc_x0, cx1 = floordiv_pos(p.x + r, 32) - o.bc_x0;
  const int cy0 = floordiv_pos(p.y - r, CELL_H) - o.bc_y0, cy1 = floordiv_pos(p.y + r, CELL_H) - o.bc_y0;
  for (int cy = cy0; cy <= cy1; cy++)
    for (int cx = cx0; cx <= cx1; cx++) {
      int2* e = ranges + o.bc_base + cy * o.bc_nx + cx;
      atomicMin(&e->x, k);
      atomicMax(&e->y, k);
    }
}
// stamps that may reach the 32 pixels xx0 .. xx0 + 31 of row yy (object frame): the window lies in one or two cells
__device__ __forceinline__ int2 brush_range(const int2* __restrict__ ranges, const ObjRec& o, int xx0, int yy) {
  const int cy = floordiv_pos(yy, CELL_H) - o.bc_y0;
  if (cy < 0 || cy >= o.bc_ny) return make_int2(0, -1);
  const int ca = floordiv_pos(xx0, 32) - o.bc_x0, cb = floordiv_pos(xx0 + 31, 32) - o.bc_x0;
  int2 rg = make_int2(INT32_MAX, -1);
  if (ca >= 0 && ca < o.bc_nx) rg = ranges[o.bc_base + cy * o.bc_nx + ca];
  if (cb != ca && cb >= 0 && cb < o.bc_nx) {
    const int2 r2 = ranges[o.bc_base + cy * o.bc_nx + cb];
    rg.x = min(rg.x, r2.x); rg.y = max(rg.y, r2.y);
  }
  return rg;
}

// ------------------------------------------------------------------------------------
// K1 cell binning for scenes of at most 1024 leaves.  One warp per cell scans the leaf objects in index
// order (= front to back), 32 at a time; ballot + popc give each overlapping object its slot, so every
// list comes out already sorted.  The warp of a cell keeps the hit masks of its 32-leaf chunks in
// registers (chunk c in lane c), takes its slice of the item pool from one global cursor and writes the
// list — no count pass, no scan.  Lists are addressed by [start, end) per cell.  The walker's heavy-first
// order is sixteen length classes (>= 48 objects ... 2, 1, 0) filled through cursors reserved once per
// block, so that the persistent walker warps take the heavy cells first and the tail of the launch is light.
// ------------------------------------------------------------------------------------
constexpr int BIN_CLASSES = 16;
constexpr int BIN_HEAVY_CLASSES = 14;   // classes of cells with at least two objects
__device__ __forceinline__ int bin_class(int n) {  // longest lists first; the last two classes: one object, none
  if (n >= 16) return n >= 48 ? 0 : n >= 40 ? 1 : n >= 32 ? 2 : n >= 28 ? 3 : n >= 24 ? 4 : n >= 20 ? 5 : 6;
  return n >= 12 ? 7 : n >= 8 ? 8 : n >= 6 ? 9 : n >= 5 ? 10 : n == 4 ? 11 : n == 3 ? 12 : n == 2 ? 13 : n == 1 ? 14 : 15;
}
// Cells whose only object is an opaque primitive covering all of them (the background rectangle: most cells of
// most frames) are finished by k_prefill when the update is a box: their pixels are that colour, and they never
// enter the walker's queue.
struct BinPrefill {
  uint32_t* fb;                     // null: off (arbitrary update shapes, continued frames, row-major order)
  uint32_t* peer_fb[7]; int n_peers;
  uint32_t* u_out;                  // optional
  int ux0, uy0, ux1, uy1;           // update box, inclusive
};
// blockDim = 256 (8 cells per block); pool slices and class positions are reserved once per block
__global__ void __launch_bounds__(256) k_bin1(const int4* __restrict__ leaf_box, const int* __restrict__ leaves, int n_leaves, Frame fr, int cell_row0,
                       int n_cells, int2* __restrict__ cell_rng, int* __restrict__ items, int* __restrict__ state /* [0] pool cursor, [1..] class counts */,
                       int* __restrict__ cls_cells /* [BIN_CLASSES][n_cells] or null */, const ObjRec* __restrict__ objs, int2* __restrict__ cell_head,
                       int* __restrict__ item_cell, int prefill_on, const int2* __restrict__ attr, int2* __restrict__ item_attr,
                       int4* __restrict__ item_rec) {
  __shared__ int s_n[8], s_c[8], s_base[8], s_pos[8];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + wid;
  const bool active = warp < n_cells;
  const int cx = active ? fr.ctx0 + warp % fr.cntx : 0, cy = active ? cell_row0 + warp / fr.cntx : 0;
  const int x0 = cx * TILE_W, x1 = x0 + TILE_W - 1, y0 = cy * CELL_H, y1 = y0 + CELL_H - 1;
  unsigned mymask = 0u;
  int n = 0, first = -1;
  if (active)
    for (int b = 0; b < n_leaves; b += 32) {
      const int li = b + lane;
      bool hit = false;
      if (li < n_leaves) {
        const int4 bb = leaf_box[li];
        hit = !(bb.x > x1 || bb.z < x0 || bb.y > y1 || bb.w < y0);
      }
      const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
      if (lane == (b >> 5)) mymask = m;
      if (n == 0 && m) first = leaves[b + __ffs((int)m) - 1];
      n += __popc(m);
    }
  int2 hd = make_int2(0, 0);
  if (lane == 0 && n == 1) {
    const ObjRec& o = objs[first];
    const int ex1 = x1 < fr.W - 1 ? x1 : fr.W - 1, ey1 = y1 < fr.H - 1 ? y1 : fr.H - 1;
    if (o.kind == K_PRIM && (o.fill.c0 >> 24) == 255u && o.pretrans < 0 && o.depth == 1 && objs[o.anc[0]].pretrans < 0 &&
        o.prim[0] + o.dx <= x0 && o.prim[2] + o.dx >= ex1 && o.prim[1] + o.dy <= y0 && o.prim[3] + o.dy >= ey1)
      hd = make_int2((int)o.fill.c0, 1 | ((objs[o.anc[0]].flags & OF_ROOT_SCENE) ? 2 : 0));
  }
  const bool prefill = prefill_on && (hd.y & 1);          // (lane 0's view): finished by k_prefill, in no class
  const int head_bit = __shfl_sync(0xFFFFFFFFu, hd.y & 1, 0);   // a background cell: its entries are never scan-converted
  if (lane == 0) { s_n[wid] = n; s_c[wid] = (active && !prefill) ? bin_class(n) : -1; }
  __syncthreads();
  if (wid == 0 && lane < 8) {
    // pool slice of the block, split by warp; class positions: one atomic per class present in the block
    int tot = 0, mine = 0;
    for (int k = 0; k < 8; k++) { if (k == lane) mine = tot; tot += s_n[k]; }
    int blk = 0;
    if (lane == 0 && tot) blk = atomicAdd(&state[0], tot);
    blk = __shfl_sync(0xFFu, blk, 0);
    s_base[lane] = blk + mine;
    const int c = s_c[lane];
    int rank = 0, same = 0, leader = lane;
    for (int k = 0; k < 8; k++) if (s_c[k] == c) { if (k < lane) rank++; same++; if (k < leader) leader = k; }
    int cb = 0;
    if (cls_cells && c >= 0 && leader == lane) cb = atomicAdd(&state[1 + c], same);
    cb = __shfl_sync(0xFFu, cb, leader);
    s_pos[lane] = cb + rank;
  }
  __syncthreads();
  if (!active) return;
  const int base = s_base[wid];
  int at = base;
  for (int b = 0; b < n_leaves; b += 32) {
    const unsigned m = __shfl_sync(0xFFFFFFFFu, mymask, b >> 5);
    if ((m >> lane) & 1u) {
      const int dst = at + __popc(m & ((1u << lane) - 1u));
      const int leaf = leaves[b + lane];
      items[dst] = leaf;
      if (item_cell) item_cell[dst] = warp;
      if (item_attr) item_attr[dst] = attr[leaf];
      if (item_rec) {
        // what k_pre_scan needs of the object, in device space, next to the list entry (two int4 per entry):
        //   {object, cell, kind | winding << 8 | background cell << 16, first row-edge slot - first listed row}
        //   paths: {first listed row, last listed row, dx, dy};  primitives: the box x0, y0, x1, y1
        const ObjRec& o = objs[leaf];
        int4 r0 = make_int4(leaf, warp, o.kind | (o.winding << 8) | (head_bit << 16), 0), r1 = make_int4(0, 0, 0, 0);
        if (o.kind == K_PRIM) r1 = make_int4(o.prim[0] + o.dx, o.prim[1] + o.dy, o.prim[2] + o.dx, o.prim[3] + o.dy);
        else if (o.kind == K_PATH) { r0.w = o.row_base - (o.ry0 + o.dy); r1 = make_int4(o.ry0 + o.dy, o.ry1 + o.dy, o.dx, o.dy); }
        item_rec[2 * (size_t)dst] = r0; item_rec[2 * (size_t)dst + 1] = r1;
      }
    }
    at += __popc(m);
  }
  if (lane == 0) {
    cell_rng[warp] = make_int2(base, base + n);
    cell_head[warp] = hd;
    if (cls_cells && s_c[wid] >= 0) cls_cells[(size_t)s_c[wid] * n_cells + s_pos[wid]] = warp;
  }
}
// The heavy-first order of k_bin1 flattened for the row compositor: position q -> {cell, list start, list end, header
// flags}; positions past the queued cells hold cell -1.  One thread per position; runs when the binning does.
// (w: bits 0-3 header flags, 4-18 tile column of the cell inside the pass, 19-30 its cell row: no division per pixel row)
__global__ void k_comp_order(const int* __restrict__ cls_cnt, const int* __restrict__ cls_cells, const int2* __restrict__ cell_rng, const int2* __restrict__ cell_head,
                             int n_cells, int4* __restrict__ order, int cntx) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_cells) return;
  int before = 0, cell = -1;
  for (int c = 0; c < BIN_CLASSES; c++) {
    const int n = cls_cnt[c];
    if (q < before + n) { cell = cls_cells[(size_t)c * n_cells + q - before]; break; }
    before += n;
  }
  int4 r = make_int4(-1, 0, 0, 0);
  if (cell >= 0) { const int2 rg = cell_rng[cell]; r = make_int4(cell, rg.x, rg.y, (cell_head[cell].y & 15) | ((cell % cntx) << 4) | ((cell / cntx) << 19)); }
  order[q] = r;
}
// The background cells of a box update (flagged by k_bin1): one warp per cell streams the cell's 16 rows of one
// colour with 128-bit stores (8 lanes per 128-byte row, 4 rows per instruction) where the cell lies inside the
// update box, pixel by pixel at the box's rim.
__global__ void __launch_bounds__(256) k_prefill(const int2* __restrict__ cell_head, Frame fr, int cell_row0, int n_cells, BinPrefill pf) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_cells) return;
  const int2 hd = cell_head[warp];
  if (!(hd.y & 1)) return;
  const int cx = fr.ctx0 + warp % fr.cntx, cy = cell_row0 + warp / fr.cntx;
  const int x0 = cx * TILE_W, y0 = cy * CELL_H;
  const uint32_t c0 = (uint32_t)hd.x;
  const bool scene_root = (hd.y & 2) != 0;
  uint32_t colmask = interval_mask32(x0, pf.ux0, pf.ux1);
  if (x0 + 31 >= fr.W) colmask &= interval_mask32(x0, 0, fr.W - 1);
  const int ya = max(max(y0, fr.band_y0), pf.uy0), yb = min(min(y0 + CELL_H - 1, fr.band_y1 - 1), pf.uy1);   // rows of the cell inside band and update
  if (pf.u_out && lane < CELL_H) {
    const int y = y0 + lane;
    if (y >= fr.band_y0 && y < fr.band_y1) pf.u_out[(size_t)y * fr.tiles_x + cx] = (scene_root || y < pf.uy0 || y > pf.uy1) ? 0u : colmask;
  }
  if (yb < ya || colmask == 0u) return;
  if (colmask == 0xFFFFFFFFu && (fr.W & 3) == 0) {
    const uint4 v = make_uint4(c0, c0, c0, c0);
    for (int y = ya + (lane >> 3); y <= yb; y += 4) {
      const size_t at = (size_t)y * fr.W + x0 + 4 * (lane & 7);
      *reinterpret_cast<uint4*>(pf.fb + at) = v;
#pragma unroll
      for (int k = 0; k < 7; k++) if (k < pf.n_peers) *reinterpret_cast<uint4*>(pf.peer_fb[k] + at) = v;
    }
    return;
  }
  if ((colmask >> lane) & 1u)
    for (int y = ya; y <= yb; y++) {
      const size_t at = (size_t)y * fr.W + x0 + lane;
      pf.fb[at] = c0;
#pragma unroll
      for (int k = 0; k < 7; k++) if (k < pf.n_peers) pf.peer_fb[k][at] = c0;
    }
}
// ------------------------------------------------------------------------------------
// K1 binning for large scenes, two levels.  Level 1 is object-parallel over coarse cells of
// COARSE x COARSE fine cells (128 x 64 pixels): one warp per leaf adds its POSITION in the leaf list to
// every coarse cell its box covers (count with fire-and-forget atomics, scan, scatter with atomic cursors),
// then k_bin_sort restores front-to-back order inside every coarse list (rank sort: positions are unique).
// Level 2 is cell-parallel like k_bin1: the warp of a fine cell scans only its coarse cell's list — two
// passes over a list of about a hundred entries instead of one scattered atomic with a return value per
// (leaf, fine cell) pair and a sort of every fine list.  Cost: O(sum of covered coarse cells + cells x
// coarse list length), not O(cells x leaves).
// ------------------------------------------------------------------------------------
constexpr int COARSE_SHIFT = 2, COARSE = 1 << COARSE_SHIFT;
template <bool FILL>
__global__ void k_bin_obj(const int4* __restrict__ leaf_box, int n_leaves, int ctiles_x, int crow0, int crow1,
                          int* __restrict__ counts, const int* __restrict__ offsets, int* __restrict__ items) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_leaves) return;
  const int4 bb = leaf_box[warp];
  constexpr int CW = 32 * COARSE, CH = CELL_H * COARSE;
  const int cx0 = max(floordiv_pos(bb.x, CW), 0), cx1 = min(floordiv_pos(bb.z, CW), ctiles_x - 1);
  const int cy0 = max(floordiv_pos(bb.y, CH), crow0), cy1 = min(floordiv_pos(bb.w, CH), crow1);
  if (cx1 < cx0 || cy1 < cy0) return;
  const int nx = cx1 - cx0 + 1, n = nx * (cy1 - cy0 + 1);
  for (int k = lane; k < n; k += 32) {
    const int cell = (cy0 + k / nx - crow0) * ctiles_x + cx0 + k % nx;
    const int pos = atomicAdd(&counts[cell], 1);
    if (FILL) items[offsets[cell] + pos] = warp;
  }
}
// Rank sort of every coarse list (ascending leaf position = front to back), one warp per list.
constexpr int SORT_SMEM = 1024;  // list entries staged in shared memory per warp
__global__ void __launch_bounds__(128) k_bin_sort(const int* __restrict__ offsets, int* __restrict__ items,
                                                  int* __restrict__ tmp, int n_cells) {
  __shared__ int s_list[4][SORT_SMEM];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (warp >= n_cells) return;
  const int a = offsets[warp], n = offsets[warp + 1] - a;
  if (n <= 1) return;
  const bool in_smem = n <= SORT_SMEM;
  int* src = in_smem ? s_list[wid] : (tmp + a);
  for (int i = lane; i < n; i += 32) src[i] = items[a + i];
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const int v = src[i];
    int rank = 0;
    for (int j = 0; j < n; j++) rank += src[j] < v;
    items[a + rank] = v;
  }
}
// Level 2: blockDim = 256 (8 fine cells per block); same outputs as k_bin1.
__global__ void __launch_bounds__(256) k_bin2(const int4* __restrict__ leaf_box, const int* __restrict__ leaves, const int* __restrict__ coarse_off,
                       const int* __restrict__ coarse_items, int ctiles_x, int crow0, Frame fr, int cell_row0, int n_cells,
                       int2* __restrict__ cell_rng, int* __restrict__ items, int* __restrict__ state, int* __restrict__ cls_cells) {
  __shared__ int s_n[8], s_c[8], s_base[8], s_pos[8];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + wid;
  const bool active = warp < n_cells;
  const int cx = active ? fr.ctx0 + warp % fr.cntx : 0, cy = active ? cell_row0 + warp / fr.cntx : 0;
  const int x0 = cx * TILE_W, x1 = x0 + TILE_W - 1, y0 = cy * CELL_H, y1 = y0 + CELL_H - 1;
  const int cc = ((cy >> COARSE_SHIFT) - crow0) * ctiles_x + (cx >> COARSE_SHIFT);
  const int la = active ? coarse_off[cc] : 0, lb = active ? coarse_off[cc + 1] : 0;
  int n = 0;
  for (int b = la; b < lb; b += 32) {
    bool hit = false;
    if (b + lane < lb) {
      const int4 bb = leaf_box[coarse_items[b + lane]];
      hit = !(bb.x > x1 || bb.z < x0 || bb.y > y1 || bb.w < y0);
    }
    n += __popc(__ballot_sync(0xFFFFFFFFu, hit));
  }
  if (lane == 0) { s_n[wid] = n; s_c[wid] = active ? bin_class(n) : -1; }
  __syncthreads();
  if (wid == 0 && lane < 8) {
    int tot = 0, mine = 0;
    for (int k = 0; k < 8; k++) { if (k == lane) mine = tot; tot += s_n[k]; }
    int blk = 0;
    if (lane == 0 && tot) blk = atomicAdd(&state[0], tot);
    blk = __shfl_sync(0xFFu, blk, 0);
    s_base[lane] = blk + mine;
    const int c = s_c[lane];
    int rank = 0, same = 0, leader = lane;
    for (int k = 0; k < 8; k++) if (s_c[k] == c) { if (k < lane) rank++; same++; if (k < leader) leader = k; }
    int cb = 0;
    if (cls_cells && c >= 0 && leader == lane) cb = atomicAdd(&state[1 + c], same);
    cb = __shfl_sync(0xFFu, cb, leader);
    s_pos[lane] = cb + rank;
  }
  __syncthreads();
  if (!active) return;
  const int base = s_base[wid];
  int at = base;
  for (int b = la; b < lb; b += 32) {
    bool hit = false;
    int li = 0;
    if (b + lane < lb) {
      li = coarse_items[b + lane];
      const int4 bb = leaf_box[li];
      hit = !(bb.x > x1 || bb.z < x0 || bb.y > y1 || bb.w < y0);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) items[at + __popc(m & ((1u << lane) - 1u))] = leaves[li];
    at += __popc(m);
  }
  if (lane == 0) {
    cell_rng[warp] = make_int2(base, base + n);
    if (cls_cells) cls_cells[(size_t)s_c[wid] * n_cells + s_pos[wid]] = warp;
  }
}

// Exclusive scan of n ints (out has n + 1 entries).  Multi-block without inter-block communication:
// block b first reduces everything in front of its 1024-element tile (coalesced, independent loads —
// n is tens of thousands, so the redundant reads are cheaper than a second launch or a look-back
// chain), then scans its tile.  The heavy-first order's 256-bin histogram is scanned on the side by
// the last warp of block 0.
__device__ __forceinline__ int block_reduce_1024(int v, int* warp_sums) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  if (lane == 0) warp_sums[wid] = v;
  __syncthreads();
  int t = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, d);
  __syncthreads();
  return t;
}
// tile_offset (optional): precomputed exclusive prefix of every 1024-element tile (large n, see
// exclusive_scan() in coherence_b200.cu); without it the block reduces its own prefix.
__global__ void k_tile_sums(const int* __restrict__ in, int* __restrict__ sums, int n) {
  __shared__ int warp_sums[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  int v = block_reduce_1024(i < n ? in[i] : 0, warp_sums);
  if (threadIdx.x == 0) sums[blockIdx.x] = v;
}
__global__ void __launch_bounds__(1024) k_exclusive_scan(const int* __restrict__ in, int* __restrict__ out, int n,
                                                         int* __restrict__ hist = nullptr,
                                                         const int* __restrict__ tile_offset = nullptr) {
  __shared__ int warp_sums[32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (hist && blockIdx.x == 0 && wid == 31) {
    int carry = 0;
    for (int base = 0; base < ORDER_BINS; base += 32) {
      int v = hist[base + lane], x = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += u; }
      hist[base + lane] = carry + x - v;
      carry += __shfl_sync(0xFFFFFFFFu, x, 31);
    }
  }
  const int tile0 = blockIdx.x * 1024;
  int before = 0;
  if (tile_offset) before = tile_offset[blockIdx.x];
  else {
    for (int i = t; i < tile0; i += 1024) before += in[i];
    before = block_reduce_1024(before, warp_sums);
  }
  const int i = tile0 + t;
  const int v = i < n ? in[i] : 0;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += u; }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int s2 = warp_sums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xFFFFFFFFu, s2, d); if (lane >= d) s2 += u; }
    warp_sums[lane] = s2;
  }
  __syncthreads();
  const int excl = before + (wid ? warp_sums[wid - 1] : 0) + x - v;
  if (i < n) out[i] = excl;
  if (i == n - 1) out[n] = excl + v;
  if (n == 0 && blockIdx.x == 0 && t == 0) out[0] = 0;
}

#include "kernels_walk.cuh"
#include "kernels_frames.cuh"
#include "kernels_shapes.cuh"
}  // namespace coh
