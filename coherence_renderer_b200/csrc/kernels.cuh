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
// most frames) are finished right here when the update is a box: their pixels are that colour, and they never
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
                       int* __restrict__ item_cell, BinPrefill pf) {
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
  const bool prefill = pf.fb != nullptr && (hd.y & 1);          // (lane 0's view)
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
      items[dst] = leaves[b + lane];
      if (item_cell) item_cell[dst] = warp;
    }
    at += __popc(m);
  }
  if (lane == 0) {
    cell_rng[warp] = make_int2(base, base + n);
    cell_head[warp] = hd;
    if (cls_cells && s_c[wid] >= 0) cls_cells[(size_t)s_c[wid] * n_cells + s_pos[wid]] = warp;
  }
  if (__shfl_sync(0xFFFFFFFFu, (int)prefill, 0)) {
    const uint32_t c0 = (uint32_t)__shfl_sync(0xFFFFFFFFu, hd.x, 0);
    const bool scene_root = (__shfl_sync(0xFFFFFFFFu, hd.y, 0) & 2) != 0;
    uint32_t colmask = interval_mask32(x0, pf.ux0, pf.ux1);
    if (x0 + 31 >= fr.W) colmask &= interval_mask32(x0, 0, fr.W - 1);
#pragma unroll 1
    for (int r = 0; r < CELL_H; r++) {
      const int y = y0 + r;
      if (y < fr.band_y0 || y >= fr.band_y1) continue;
      const uint32_t u = (y >= pf.uy0 && y <= pf.uy1) ? colmask : 0u;
      if (pf.u_out && lane == 0) pf.u_out[(size_t)y * fr.tiles_x + cx] = scene_root ? 0u : u;
      if ((u >> lane) & 1u) {
        const size_t at = (size_t)y * fr.W + x0 + lane;
        pf.fb[at] = c0;
        for (int k = 0; k < pf.n_peers; k++) pf.peer_fb[k][at] = c0;
      }
    }
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
  __shared__ int s_prefix[32 * 33];
  __shared__ uint32_t s_aa[WALK_WARPS][32 * AA_WORDS];
  __shared__ uint32_t s_acc[WALK_WARPS][WALK_H][32];
  __shared__ StagedEdge s_stage[WALK_WARPS][32];
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) s_prefix[i] = (&P.aa->prefix[0][0])[i];
  __syncthreads();
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

// ------------------------------------------------------------------------------------
// K2 stand-alone (export path): one thread per pixel row of one edge list writes the
// shape and coverage bit-rows into global bit-frames of `nw` words per row.
// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// Three-phase frames for scenes of plain-filled paths and primitives (the lion).  The fused walker's
// longest work item is a chain of antialiasing calls that must run one after the other; here scan
// conversion and antialiasing of every (cell item, row) pair are independent work for the whole GPU,
// and the front-to-back walk that remains only composites:
//   k_pre_scan  thread / pair: shape and coverage words (the walker's scan phase, for every candidate)
//   k_pre_vis   lane / (cell, row): which edge pixels can still show — `u` pruned by the minshapes of the
//               opaque objects in front (a superset of the exact `u`: edge pixels never count as covered
//               here), and the list of pairs that need antialiasing
//   k_pre_aa    warp / listed pair: opacity bytes (aa_tile)
//   k_walk<PRE> composite with the exact `u`; every pixel it antialiases is in the superset.
// ------------------------------------------------------------------------------------
__global__ void k_pre_scan(WalkParams P, int n_pairs, uint2* __restrict__ sc) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_pairs) return;
  const int item = pair / CELL_H, row = pair % CELL_H;
  const int cell = P.item_cell[item];
  if (P.cell_head[cell].y & 1) return;   // a background cell: finished by the binning kernel or the walker's fast path, nobody reads these words
  const int tile = P.fr.ctx0 + cell % P.fr.cntx, by = cell / P.fr.cntx;
  const int tx0 = tile * TILE_W, my_y = (P.cell_row0 + by) * CELL_H + row;
  const ObjRec& o = P.objs[P.cell_items[item]];
  uint32_t S = 0u, C = 0u;
  if (my_y >= P.fr.band_y0 && my_y < P.fr.band_y1 && !(o.by0 > my_y || o.by1 < my_y || o.bx0 > tx0 + 31 || o.bx1 < tx0)) {
    const int yy = my_y - o.dy, xx0 = tx0 - o.dx;
    if (o.kind == K_PRIM) {
      if (yy >= o.prim[1] && yy <= o.prim[3]) S = interval_mask32(xx0, o.prim[0], o.prim[2]);
    } else if (yy >= o.ry0 && yy <= o.ry1) {
      const int slot = o.row_base + yy - o.ry0;
      const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
      bool ok = true;
      const uint2 w = scan_row_word(P.edges, P.rowedge_idx + a, b - a, yy, o.winding, xx0, ok);
      if (!ok) *P.error_flag = 1;
      S = w.x; C = w.y;
    }
  }
  sc[pair] = make_uint2(S, C);
}
// blockDim = 128: 8 (cell, 16 rows) groups per block
__global__ void k_pre_vis(WalkParams P, const uint2* __restrict__ sc, int4* __restrict__ list /* pair, object, edge mask, (tile << 16 | row of the frame) */, int* __restrict__ list_n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int cell = t / CELL_H, row = t % CELL_H;
  if (cell >= P.n_cells || (P.cell_head[cell].y & 1)) return;
  const int tile = P.fr.ctx0 + cell % P.fr.cntx, by = cell / P.fr.cntx;
  const int tx0 = tile * TILE_W, my_y = (P.cell_row0 + by) * CELL_H + row;
  uint32_t u = 0u;
  if (my_y >= P.fr.band_y0 && my_y < P.fr.band_y1) {
    if (P.u_init) u = P.u_init[(size_t)my_y * P.fr.tiles_x + tile];
    else u = (my_y >= P.uy0 && my_y <= P.uy1) ? interval_mask32(tx0, P.ux0, P.ux1) : 0u;
    if (tx0 + 31 >= P.fr.W) u &= interval_mask32(tx0, 0, P.fr.W - 1);
  }
  const int2 cell_rg = P.cell_rng[cell];
  const int it0 = cell_rg.x, it1 = cell_rg.y;
  for (int it = it0; it < it1; it++) {
    const int oi = P.cell_items[it];
    const ObjRec& o = P.objs[oi];
    const size_t pair = (size_t)it * CELL_H + row;
    const uint2 w = sc[pair];
    const uint32_t M = w.x & ~w.y;
    const uint32_t e = (o.kind == K_PATH) ? (w.x & ~M & u) : 0u;
    if (e) list[atomicAdd(list_n, 1)] = make_int4((int)pair, oi, (int)e, (tile << 16) | my_y);
    if (o.flags & OF_OCCLUDES) u &= ~M;   // opaque fill, no dissolve on the way up: its interior hides what is behind
  }
}
__global__ void __launch_bounds__(256) k_pre_aa(WalkParams P, const int4* __restrict__ list, const int* __restrict__ list_n, uint8_t* __restrict__ op) {
  __shared__ int s_prefix[32 * 33];
  __shared__ uint32_t s_aa[8][32 * AA_WORDS];
  __shared__ StagedEdge s_stage[8][32];
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) s_prefix[i] = (&P.aa->prefix[0][0])[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = *list_n, n_warps = gridDim.x * 8;
  for (int i = blockIdx.x * 8 + wid; i < n; i += n_warps) {
    const int4 ent = list[i];
    const ObjRec& o = P.objs[ent.y];
    const int yy = (ent.w & 0xFFFF) - o.dy, xx0 = (ent.w >> 16) * TILE_W - o.dx;
    const int slot = o.row_base + yy - o.ry0;
    const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
    bool ok;
    const int v = aa_tile(P.edges, P.rowedge_idx + a, b - a, o.aa_winding, xx0, yy, (uint32_t)ent.z, s_aa[wid], s_stage[wid], s_prefix, P.aa->volume, lane, ok);
    if (!ok) *P.error_flag = 1;
    op[(size_t)ent.x * 32 + lane] = (uint8_t)v;
    __syncwarp();
  }
}

constexpr int SCAN_CHUNK_WORDS = 8;  // one thread scans a 256-pixel window of one row
__global__ void k_scan_rows(const EdgeRec* __restrict__ edges, int n_edges, int winding, int y0, int n_rows,
                            int wx0, int nw, uint32_t* __restrict__ S, uint32_t* __restrict__ C, int* error_flag) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  int w0 = blockIdx.y * SCAN_CHUNK_WORDS;
  if (r >= n_rows || w0 >= nw) return;
  SinkMem sink;
  sink.wx0 = wx0 + 32 * w0; sink.nwords = min(SCAN_CHUNK_WORDS, nw - w0); sink.stride = 1;
  sink.S = S + (size_t)r * nw + w0; sink.C = C + (size_t)r * nw + w0;
  if (!scan_row(edges, nullptr, n_edges, 1, y0 + r, winding, false, sink.wx0, sink.wx0 + 32 * sink.nwords - 1, sink))
    *error_flag = 1;
}

// ------------------------------------------------------------------------------------
// K4 stand-alone (export path): AA opacity for every set pixel of a bit-frame `Q`
// (rows y0.., nw words per row starting at pixel wx0).  One warp per (row, word).
// Output: dense bytes out[r][nw*32].
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_aa_rows(const EdgeRec* __restrict__ edges, int n_edges, int winding,
                                                 const uint32_t* __restrict__ Q, int y0, int n_rows, int wx0, int nw,
                                                 const AATable* __restrict__ aa, uint8_t* __restrict__ out, int* error_flag) {
  __shared__ int s_prefix[32 * 33];
  __shared__ uint32_t s_aa[8][32 * AA_WORDS];
  __shared__ StagedEdge s_stage[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int w = blockIdx.x * 8 + wid, r = blockIdx.y;
  const uint32_t q = (w < nw && r < n_rows) ? Q[(size_t)r * nw + w] : 0u;
  if (!__syncthreads_or(q != 0u)) return;   // most blocks of a sparse frame have nothing to sample: leave before loading the table
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) s_prefix[i] = (&aa->prefix[0][0])[i];
  __syncthreads();
  if (!q) return;
  bool ok;
  int op = aa_tile(edges, nullptr, n_edges, winding, wx0 + 32 * w, y0 + r, q, s_aa[wid], s_stage[wid], s_prefix, aa->volume, lane, ok);
  if (!ok) *error_flag = 1;
  if ((q >> lane) & 1u) out[((size_t)r * nw + w) * 32 + lane] = (uint8_t)op;
}

// ------------------------------------------------------------------------------------
// K3: span sets <-> bit-frames and word-wise set algebra.
// A device span set is CSR: rows y0 .. y0+n_rows-1, row_ptr[n_rows+1], spans (x, len).
// ------------------------------------------------------------------------------------
__global__ void k_spans_to_bits(const int* __restrict__ row_ptr, const int2* __restrict__ spans, int src_y0,
                                int src_rows, int y0, int n_rows, int wx0, int nw, uint32_t* __restrict__ bits) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;  // destination row
  if (r >= n_rows) return;
  int sr = y0 + r - src_y0;
  if (sr < 0 || sr >= src_rows) return;
  uint32_t* row = bits + (size_t)r * nw;
  for (int k = row_ptr[sr]; k < row_ptr[sr + 1]; k++) {
    int2 s = spans[k];
    or_interval(row, 1, nw, wx0, s.x, s.x + s.y - 1);
  }
}
// Brush.shape_of_brushstroke (brush.ml:135-173): the union of the (2r+1)^2 boxes around the stamp centres.
// One thread per (stamp, row of its box); rows y0 .., nw words per row starting at pixel wx0.
__global__ void k_stamp_boxes_to_bits(const int2* __restrict__ points, int n_points, int r, int y0, int n_rows, int wx0, int nw,
                                      uint32_t* __restrict__ bits) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, side = 2 * r + 1;
  if (t >= n_points * side) return;
  const int2 p = points[t / side];
  const int row = p.y - r + t % side - y0;
  if (row < 0 || row >= n_rows) return;
  int a = p.x - r - wx0, b = p.x + r - wx0;
  if (b < 0 || a >= nw * 32) return;
  a = max(a, 0); b = min(b, nw * 32 - 1);
  uint32_t* rowp = bits + (size_t)row * nw;
  for (int w = a >> 5; w <= (b >> 5); w++) {
    const int lo = max(a, w * 32) & 31, hi = min(b, w * 32 + 31) & 31;
    atomicOr(&rowp[w], (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31 - hi)));
  }
}
// op: 0 OR, 1 ANDNOT (a & ~b), 2 AND
__global__ void k_bitop(const uint32_t* a, const uint32_t* b, uint32_t* out,
                        size_t n, int op) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = a[i], y = b[i];
  out[i] = op == 0 ? (x | y) : op == 1 ? (x & ~y) : (x & y);
}
// Dilation by (2m+1) x (2n+1) (Sprite.bloat, sprite.ml:1749-1864) on a bit-frame that
// already has a margin of m pixels / n rows around the set.  One thread per word.
__global__ void k_dilate(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n_rows, int nw, int m, int n) {
  int w = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (w >= nw || r >= n_rows) return;
  uint32_t acc = 0u;
  for (int rr = max(0, r - n); rr <= min(n_rows - 1, r + n); rr++) {
    const uint32_t* row = in + (size_t)rr * nw;
    // OR of the row shifted by -m..+m pixels, gathered for this word
    for (int s = -m; s <= m; s++) {
      // bit i of result word w comes from pixel 32w + i - s
      int q = 32 * w - s;            // source pixel of bit 0
      int qw = q >> 5, qb = q & 31;  // arithmetic shift: floor
      uint32_t lo = (qw >= 0 && qw < nw) ? row[qw] : 0u;
      uint32_t hi = (qw + 1 >= 0 && qw + 1 < nw) ? row[qw + 1] : 0u;
      acc |= qb ? ((lo >> qb) | (hi << (32 - qb))) : lo;
    }
  }
  out[(size_t)r * nw + w] = acc;
}
// Run extraction: count the maximal runs of every row (thread per row), then fill.
__global__ void k_count_runs(const uint32_t* __restrict__ bits, int n_rows, int nw, int* __restrict__ counts,
                             unsigned long long* __restrict__ card) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint32_t* row = bits + (size_t)r * nw;
  int n = 0; uint32_t carry = 0u; unsigned long long px = 0;
  for (int w = 0; w < nw; w++) {
    uint32_t v = row[w];
    n += __popc(v & ~((v << 1) | carry));
    px += __popc(v);
    carry = v >> 31;
  }
  counts[r] = n;
  if (card && px) atomicAdd(card, px);
}
__global__ void k_fill_runs(const uint32_t* __restrict__ bits, int n_rows, int nw, int wx0,
                            const int* __restrict__ row_ptr, int2* __restrict__ spans) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint32_t* row = bits + (size_t)r * nw;
  int k = row_ptr[r];
  int start = 0; bool in = false;
  for (int w = 0; w < nw; w++) {
    uint32_t v = row[w];
    int base = wx0 + 32 * w;
    int pos = 0;
    while (pos < 32) {
      if (!in) {
        uint32_t rest = pos ? (v >> pos) : v;
        if (!rest) break;
        pos += __ffs((int)rest) - 1;
        start = base + pos; in = true;
      } else {
        uint32_t rest = ~(pos ? (v >> pos) : v);
        if (pos) rest &= (0xFFFFFFFFu >> pos);  // bits shifted in from above are not pixels
        if (!rest) { pos = 32; break; }
        pos += __ffs((int)rest) - 1;
        spans[k++] = make_int2(start, base + pos - start);
        in = false;
      }
    }
  }
  if (in) spans[k++] = make_int2(start, wx0 + 32 * nw - start);
}
// Sprite.translate_shape (sprite.ml:470-484) on a device span set: spans move by dx (rows move by
// changing y0 on the host side).
__global__ void k_translate_spans(const int2* __restrict__ in, int2* __restrict__ out, int n, int dx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { int2 s = in[i]; out[i] = make_int2(s.x + dx, s.y); }
}
// Per-object alias offsets changed in place (Render.translate_renderobject -> Cache.addtranslation):
// shift the device-space boxes the binning reads.  delta = new offset - old offset.
__global__ void k_move_leaves(ObjRec* __restrict__ objs, int4* __restrict__ leaf_box, const int* __restrict__ leaves,
                              int n_leaves, int first_obj, int last_obj, int ddx, int ddy) {
  int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= n_leaves) return;
  int oi = leaves[li];
  if (oi < first_obj || oi > last_obj) return;
  ObjRec& o = objs[oi];
  o.dx += ddx; o.dy += ddy; o.bx0 += ddx; o.bx1 += ddx; o.by0 += ddy; o.by1 += ddy;
  leaf_box[li] = make_int4(o.bx0, o.by0, o.bx1, o.by1);
}
// ------------------------------------------------------------------------------------
// Filters (render.ml:1080-1131, 1248-1265; filters.ml).  Frame-sized RGBA8 canvases and bit-frames
// (nw words per row, bit 0 of word 0 = pixel x 0).
// ------------------------------------------------------------------------------------
// canvas[p] = clear for every pixel p of the bit-frame
__global__ void k_clear_in_bits(uint32_t* __restrict__ canvas, const uint32_t* __restrict__ bits, int W, int H, int nw) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  if ((bits[(size_t)y * nw + (x >> 5)] >> (x & 31)) & 1u) canvas[(size_t)y * W + x] = 0u;
}
// Filters.monochrome: sprite_map Colour.monochrome (colour.ml: average of r, g, b; alpha kept)
__global__ void k_monochrome(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = in[i];
  const uint32_t av = ((c & 255u) + ((c >> 8) & 255u) + ((c >> 16) & 255u)) / 3u;
  out[i] = av | (av << 8) | (av << 16) | (c & 0xFF000000u);
}
// The filter geometry's matte inside T (render.ml:1099): alpha of `dissolve fill opacity` with the
// antialiased opacity bytes `op` (Polygon.polygon_sprite samples every pixel it is given, minshape
// pixels included); `finished` = its opaque pixels (1100-1103).  One word per warp.
__global__ void k_filter_matte(const uint32_t* __restrict__ T, const uint8_t* __restrict__ op,
                               uint32_t colour, int W, int H, int nw, uint8_t* __restrict__ alpha, uint32_t* __restrict__ finished) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), y = blockIdx.y, lane = threadIdx.x & 31;
  if (w >= nw || y >= H) return;
  const uint32_t t = T[(size_t)y * nw + w];
  const int x = 32 * w + lane;
  int a = 0;
  if (((t >> lane) & 1u) && x < W) {
    a = (int)(px_dissolve(colour, op[(size_t)y * nw * 32 + x]) >> 24);
    alpha[(size_t)y * W + x] = (uint8_t)a;
  }
  const uint32_t f = __ballot_sync(0xFFFFFFFFu, a == 255);
  if (lane == 0) finished[(size_t)y * nw + w] = f & t;
}
// blend' (render.ml:1248-1265) and the composite of the filter's sprite into the accumulator
// (render.ml:1290-1291): fb = over fb (pd_plus (dissolve Z (255 - alpha)) (dissolve Y alpha)) on T.
// Pixels a scene did not render are clear in Z / Y, which both operators treat as absent.
__global__ void k_filter_blend(const uint32_t* __restrict__ T, const uint8_t* __restrict__ alpha, const uint32_t* __restrict__ Z,
                               const uint32_t* __restrict__ Y, uint32_t* __restrict__ fb, int W, int H, int nw) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  if (!((T[(size_t)y * nw + (x >> 5)] >> (x & 31)) & 1u)) return;
  const size_t i = (size_t)y * W + x;
  const int a = alpha[i];
  const uint32_t z = px_dissolve(Z[i], 255 - a), yy = Y ? px_dissolve(Y[i], a) : 0u;
  fb[i] = px_over(fb[i], px_plus(z, yy));
}
// update & ~opaque(fb): where the background list is still visible under the scene pass
__global__ void k_not_opaque_bits(const uint32_t* __restrict__ fb, const uint32_t* __restrict__ U, uint32_t* __restrict__ out, int W, int H, int nw) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), y = blockIdx.y, lane = threadIdx.x & 31;
  if (w >= nw || y >= H) return;
  const int x = 32 * w + lane;
  const bool opq = x < W && (fb[(size_t)y * W + x] >> 24) == 255u;
  const uint32_t o = __ballot_sync(0xFFFFFFFFu, opq);
  if (lane == 0) out[(size_t)y * nw + w] = U[(size_t)y * nw + w] & ~o;
}

// ------------------------------------------------------------------------------------
// K6 convolve (convolve.ml:115-232) on dense RGBA8 canvases [h][w]; pixels outside the canvas
// read as clear, like the 2r border of Sprite.flatten_sprite (convolve.ml:247).  One pass per
// launch (horizontal, then vertical on the re-quantised result): integer sums, truncating
// division, r,g clamped to alpha for XY kernels (the blue clamp of the reference is a no-op,
// convolve.ml:118), plain division for the unit kernel (161-204).
// ------------------------------------------------------------------------------------
__global__ void k_conv_pass(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int w, int h, int r,
                            int kind /*1 unit, 2 xy*/, const int* __restrict__ taps, int total, int vertical) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  int tr = 0, tg = 0, tb = 0, ta = 0;
  for (int q = -r; q <= r; q++) {
    int xx = vertical ? x : x + q, yy = vertical ? y + q : y;
    uint32_t c = (xx >= 0 && xx < w && yy >= 0 && yy < h) ? in[(size_t)yy * w + xx] : 0u;
    int k = kind == 2 ? taps[q + r] : 1;
    tr += (int)(c & 255u) * k; tg += (int)((c >> 8) & 255u) * k; tb += (int)((c >> 16) & 255u) * k; ta += (int)(c >> 24) * k;
  }
  int d = kind == 2 ? total : (2 * r + 1);
  tr /= d; tg /= d; tb /= d; ta /= d;
  if (kind == 2) { tr = min(ta, tr); tg = min(ta, tg); }
  out[(size_t)y * w + x] = (uint32_t)tr | ((uint32_t)tg << 8) | ((uint32_t)tb << 16) | ((uint32_t)ta << 24);
}
// AA raster of a plain-filled polygon from dense opacity bytes: dissolve fill opacity (polygon.ml:733-738)
__global__ void k_raster_plain(const uint8_t* __restrict__ opacity, uint32_t* __restrict__ out, size_t n, uint32_t colour) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = px_dissolve(colour, opacity[i]);
}
__global__ void k_fill_words(uint32_t* __restrict__ p, size_t n, uint32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// 32 bits of a bit-row starting at an arbitrary bit offset (zeros outside the row)
__device__ __forceinline__ uint32_t load_bits32(const uint32_t* __restrict__ row, int nw, int bitoff) {
  const int qw = bitoff >> 5, qb = bitoff & 31;
  const uint32_t lo = (qw >= 0 && qw < nw) ? row[qw] : 0u;
  const uint32_t hi = (qw + 1 >= 0 && qw + 1 < nw) ? row[qw + 1] : 0u;
  return qb ? ((lo >> qb) | (hi << (32 - qb))) : lo;
}
// Box-shaped bit-frame (Sprite.box) or clear.
__global__ void k_fill_box_bits(uint32_t* __restrict__ bits, int n_rows, int nw, int wx0, int y0, int bx0, int by0,
                                int bx1, int by1) {
  int w = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (w >= nw || r >= n_rows) return;
  int y = y0 + r;
  bits[(size_t)r * nw + w] = (y >= by0 && y <= by1) ? interval_mask32(wx0 + 32 * w, bx0, bx1) : 0u;
}
// RGB888 export of a framebuffer rectangle (wxgui.ml:417-424 plot_sprite byte layout).
__global__ void k_rgb888(const uint32_t* __restrict__ fb, int W, int x0, int y0, int w, int h, uint8_t* __restrict__ out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  uint32_t c = fb[(size_t)(y0 + y) * W + x0 + x];
  uint8_t* p = out + ((size_t)y * w + x) * 3;
  p[0] = c & 255u; p[1] = (c >> 8) & 255u; p[2] = (c >> 16) & 255u;
}
// Scatter per-pixel values given in canonical span order into a dense canvas: thread per row.
template <class T>
__global__ void k_scatter_spans(const int* __restrict__ row_ptr, const int2* __restrict__ spans,
                                const long long* __restrict__ px_off, int n_rows, int row0, int wx0, int w,
                                const T* __restrict__ in, T* __restrict__ dense) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  long long o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    int2 s = spans[k];
    for (int i = 0; i < s.y; i++) dense[(size_t)(row0 + r) * w + (s.x + i - wx0)] = in[o++];
  }
}
// Gather dense per-pixel values in canonical span order: thread per row.
template <class T>
__global__ void k_gather_spans(const int* __restrict__ row_ptr, const int2* __restrict__ spans,
                               const long long* __restrict__ px_off, int n_rows, int wx0, int pitch,
                               const T* __restrict__ dense, T* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  long long o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    int2 s = spans[k];
    for (int i = 0; i < s.y; i++) out[o++] = dense[(size_t)r * pitch + (s.x + i - wx0)];
  }
}

}  // namespace coh
