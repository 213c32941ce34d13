// kernels_frames.cuh — part of kernels.cuh (included inside namespace coh, in order): stand-alone scan / AA kernels and the three-phase frame kernels (k_pre_scan, k_pre_vis, k_pre_aa).

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
__global__ void __launch_bounds__(128) k_pre_scan(WalkParams P, int n_pairs, uint2* __restrict__ sc, const int4* __restrict__ item_rec, int* __restrict__ counters) {
  cudaTriggerProgrammaticLaunchCompletion();   // (PDL: the visibility kernel's blocks may be scheduled as this grid's last blocks start)
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair < 2) counters[pair] = 0;   // list length and next pair of the kernels behind (they wait for this grid): no memset node in front of the frame
  if (pair >= n_pairs) return;
  const int item = pair / CELL_H, row = pair % CELL_H;
  // the list entry carries what is needed of the object (written by k_bin1): no walk through ObjRec for paths and primitives
  const int4 r0 = item_rec[2 * (size_t)item], r1 = item_rec[2 * (size_t)item + 1];
  const int cell = r0.y, kind = r0.z & 255;
  if (!P.resume && ((r0.z >> 16) & 1)) return;   // a background cell: finished by k_prefill or the walker's fast path, nobody reads these words (a continued frame has no such shortcut)
  const int tile = P.fr.ctx0 + cell % P.fr.cntx, by = cell / P.fr.cntx;
  const int tx0 = tile * TILE_W, my_y = (P.cell_row0 + by) * CELL_H + row;
  uint32_t S = 0u, C = 0u;
  const bool in_band = my_y >= P.fr.band_y0 && my_y < P.fr.band_y1;
  if (in_band && kind == K_PRIM) {
    if (my_y >= r1.y && my_y <= r1.w) S = interval_mask32(tx0, r1.x, r1.z);
  } else if (in_band && kind == K_PATH) {
    if (my_y >= r1.x && my_y <= r1.y) {
      const int slot = r0.w + my_y;
      const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
      bool ok = true;
      const uint2 w = scan_row_word(P.edges, P.rowedge_idx + a, b - a, my_y - r1.w, (r0.z >> 8) & 255, tx0 - r1.z, ok);
      if (!ok) *P.error_flag = 1;
      S = w.x; C = w.y;
    }
  } else if (in_band) {
    const ObjRec& o = P.objs[r0.x];
    if (!(o.by0 > my_y || o.by1 < my_y || o.bx0 > tx0 + 31 || o.bx1 < tx0)) {
    const int yy = my_y - o.dy, xx0 = tx0 - o.dx;
    if (o.kind == K_PRIM) {
      if (yy >= o.prim[1] && yy <= o.prim[3]) S = interval_mask32(xx0, o.prim[0], o.prim[2]);
    } else if (o.kind == K_CONV) {   // Convolved object: shape / minshape kept as bit-rows by the scene
      if (yy >= o.cv_y0 && yy < o.cv_y0 + o.cv_h) {
        const uint32_t* rowS = P.conv_bits + o.cv_bits + (size_t)(yy - o.cv_y0) * o.cv_nw;
        S = conv_load_bits32(rowS, o.cv_nw, xx0 - o.cv_x0);
        C = S & ~conv_load_bits32(rowS + (size_t)o.cv_h * o.cv_nw, o.cv_nw, xx0 - o.cv_x0);
      }
    } else if (yy >= o.ry0 && yy <= o.ry1) {
      const int slot = o.row_base + yy - o.ry0;
      const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
      bool ok = true;
      const uint2 w = scan_row_word(P.edges, P.rowedge_idx + a, b - a, yy, o.winding, xx0, ok);
      if (!ok) *P.error_flag = 1;
      S = w.x; C = w.y;
    }
    }
  }
  sc[pair] = make_uint2(S, C);
}
// blockDim = 128: 8 (cell, 16 rows) groups per block
__global__ void k_pre_vis(WalkParams P, const uint2* __restrict__ sc, int4* __restrict__ list /* pair, object, edge mask, (tile << 16 | row of the frame) */, int* __restrict__ list_n,
                          const int2* __restrict__ item_attr) {
  cudaTriggerProgrammaticLaunchCompletion();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int cell = t / CELL_H, row = t % CELL_H;
  if (cell >= P.n_cells || (!P.resume && (P.cell_head[cell].y & 1))) return;   // (cell_head: written when the cells were binned)
  cudaGridDependencySynchronize();   // the scan words of k_pre_scan
  asm volatile("" : "+l"(sc));         // (`sc` is const __restrict__: nothing read through it may be hoisted above the wait)
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
  // the entries' attributes ride along with the lists (k_bin1): is-path, occludes — four entries are fetched at a
  // time so that their loads are in flight together; only u chains one entry to the next
  for (int it = it0; it < it1 && u != 0u; it += 4) {
    int2 at[4]; uint2 w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const bool in = it + k < it1;
      at[k] = in ? item_attr[it + k] : make_int2(0, 0);
      w[k] = in ? sc[(size_t)(it + k) * CELL_H + row] : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t M = w[k].x & ~w[k].y;
      const uint32_t e = (at[k].y & 1) ? (w[k].x & w[k].y & u) : 0u;   // shape - minshape, still uncovered
      if (e) {
        // one atomic per group of lanes that arrive here together
        const uint32_t m = __activemask();
        const int leader = __ffs(m) - 1, lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(list_n, __popc(m));
        base = __shfl_sync(m, base, leader);
        list[base + __popc(m & ((1u << lane) - 1u))] = make_int4((it + k) * CELL_H + row, P.cell_items[it + k], (int)e, (tile << 16) | my_y);
      }
      if (at[k].y & 4) u &= ~M;   // opaque fill, no dissolve on the way up: its interior hides what is behind
    }
  }
}
// General antialiasing of listed pairs (bit-rows in shared memory, any number of crossings): the reference form of
// k_pre_aa_runs (option "aa_general").
__global__ void __launch_bounds__(256) k_pre_aa(WalkParams P, const int4* __restrict__ list, const int* __restrict__ list_n, uint8_t* __restrict__ op) {
  __shared__ int s_prefix[32 * 33];
  __shared__ uint32_t s_aa[8][32 * AA_WORDS];
  __shared__ StagedEdge s_stage[8][32];
  const int n = *list_n;
  if (blockIdx.x * 8 >= n) return;   // (the grid is sized before the list length is known)
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) s_prefix[i] = (&P.aa->prefix[0][0])[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_warps = gridDim.x * 8;
  for (int i = blockIdx.x * 8 + wid; i < n; i += n_warps) {
    const int4 ent = list[i];
    const uint32_t edge = (uint32_t)ent.z;
    const ObjRec& o = P.objs[ent.y];
    const int yy = (ent.w & 0xFFFF) - o.dy, xx0 = (ent.w >> 16) * TILE_W - o.dx;
    const int slot = o.row_base + yy - o.ry0;
    const int a = P.rowedge_ptr[slot], b = P.rowedge_ptr[slot + 1];
    bool ok;
    const int v = aa_tile(P.edges, P.rowedge_idx + a, b - a, o.aa_winding, xx0, yy, edge, s_aa[wid], s_stage[wid], s_prefix, P.aa->volume, lane, ok);
    if (!ok) *P.error_flag = 1;
    if ((edge >> lane) & 1u) op[(size_t)ent.x * 32 + lane] = (uint8_t)v;
    __syncwarp();
  }
}
// Antialiasing of listed pairs in interval form (raster_core.cuh, AaScan): lane j evaluates super-sampled row
// 16 y - 32 + j of the x16 edge list as a run [lo, hi] (minus at most one gap) of the columns under the pair's edge
// pixels — no bit-row, no crossing lists — and every edge pixel is two prefix-table reads per lane plus one
// redux.sync for two pixels.  A pair with a row that is not provably of that form (2 % on the lion) is retried
// run by run of its edge pixels (narrower windows); what is still complex (a few dozen pairs of a frame) takes the
// general bit-row routine right here — a kernel of its own for them costs a launch and the latency of one pair.
constexpr int AA2_WARPS = 8;
// the edge pixels `edge` (a subset of the pair's) against the staged candidates; returns false when some row is complex
__device__ __forceinline__ bool aa_runs_pixels(const EdgeRec* __restrict__ edges, const int* __restrict__ idx, int n_cand, int winding,
                                               int xx0, int yy, uint32_t edge, AaEdge* __restrict__ stage, const int* __restrict__ prow,
                                               int lane, int& mytot) {
  const int wlo = 16 * xx0 - 32;
  const int nlo = wlo + 16 * (__ffs((int)edge) - 1), nhi = wlo + 16 * (31 - __clz((int)edge)) + 31;
  AaScan sc; sc.begin(16 * yy - 32 + lane, nlo, nhi);
  for (int base = 0; base < n_cand; base += 32) {
    if (base + lane < n_cand) stage[lane] = make_aa_edge(edges[idx[base + lane]], nlo, nhi);
    __syncwarp();
    const int cnt = min(32, n_cand - base);
#pragma unroll 1
    for (int k = 0; k < cnt; k++) sc.edge(stage[k]);
    __syncwarp();
  }
  int lo, hi, glo, ghi;
  const bool simple = sc.finish(winding, lo, hi, glo, ghi);
  if (!__all_sync(0xFFFFFFFFu, simple)) return false;
  const bool any_gap = __any_sync(0xFFFFFFFFu, glo <= ghi);   // some row is two runs: the hull minus a gap
  uint32_t e = edge;
  // run by run of edge pixels, two pixels per warp reduction: a table sum is below 2^16, so two pixels share
  // one register through redux.sync
  while (e) {
    const int b0 = __ffs((int)e) - 1;
    const uint32_t tz = ~(e >> b0);
    const int len = min(tz ? (__ffs((int)tz) - 1) : 32, 32 - b0);
    e &= ~((len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << b0);
#pragma unroll 1
    for (int k = 0; k < len; k += 2) {
      const int w0 = wlo + 16 * (b0 + k);
      const bool two = k + 1 < len;
      int p = aa_interval_sum(prow, lo, hi, w0);
      int q = two ? aa_interval_sum(prow, lo, hi, w0 + 16) : 0;
      if (any_gap) {
        p -= aa_interval_sum(prow, glo, ghi, w0);
        q -= two ? aa_interval_sum(prow, glo, ghi, w0 + 16) : 0;
      }
      const unsigned tot = __reduce_add_sync(0xFFFFFFFFu, (unsigned)p | ((unsigned)q << 16));
      mytot = (lane == b0 + k) ? (int)(tot & 0xFFFFu) : mytot;
      mytot = (lane == b0 + k + 1) ? (int)(tot >> 16) : mytot;
    }
  }
  return true;
}
static_assert(sizeof(AaEdge) >= sizeof(StagedEdge), "the stage of the interval scan doubles as the stage of the general scan");
__global__ void __launch_bounds__(AA2_WARPS * 32, 4) k_pre_aa_runs(WalkParams P, const int4* __restrict__ list, const int* __restrict__ list_n,
                                                                  uint8_t* __restrict__ op, int* __restrict__ next) {
  __shared__ __align__(16) int s_prefix[32 * 33];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ AaEdge s_stage[AA2_WARPS][32];
  __shared__ uint32_t s_aa[AA2_WARPS][32 * AA_WORDS];   // bit-rows of the general routine (rarely touched)
  cudaTriggerProgrammaticLaunchCompletion();
  stage_aa_table(s_prefix, &s_bar, P.aa);   // (the table is a constant of the context: staged while k_pre_vis finishes)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  cudaGridDependencySynchronize();   // the list of k_pre_vis
  asm volatile("" : "+l"(list), "+l"(list_n));   // (const __restrict__ pointers: a load through them is invariant to the compiler and
  const int n = *list_n;                          //  may otherwise be hoisted above the wait — the list length was)
  const int* prow = s_prefix + lane * 33;
  AaEdge* stage = s_stage[wid];
  // one resident wave of warps; pairs come off a counter (their cost varies with the number of candidate edges)
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(next, 1);
    i = __shfl_sync(0xFFFFFFFFu, i, 0);
    if (i >= n) break;
    const int4 ent = list[i];
    const ObjRec& o = P.objs[ent.y];
    const int yy = (ent.w & 0xFFFF) - o.dy, xx0 = (ent.w >> 16) * TILE_W - o.dx;
    const int slot = o.row_base + yy - o.ry0;
    const int ea = P.rowedge_ptr[slot], n_cand = P.rowedge_ptr[slot + 1] - ea;
    const int winding = o.aa_winding;
    const uint32_t edge = (uint32_t)ent.z;
    int mytot = 0;
    uint32_t mine = edge;   // the pixels this kernel finishes
    if (!aa_runs_pixels(P.edges, P.rowedge_idx + ea, n_cand, winding, xx0, yy, edge, stage, prow, lane, mytot)) {
      // retry run by run of the edge pixels: each run sees only the part of the object near it
      uint32_t e = edge, hard = 0u;
      while (e) {
        const int b0 = __ffs((int)e) - 1;
        const uint32_t tz = ~(e >> b0);
        const int len = min(tz ? (__ffs((int)tz) - 1) : 32, 32 - b0);
        const uint32_t run = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << b0;
        e &= ~run;
        if (run == edge || !aa_runs_pixels(P.edges, P.rowedge_idx + ea, n_cand, winding, xx0, yy, run, stage, prow, lane, mytot)) hard |= run;
      }
      if (hard) {
        bool ok;
        const int v = aa_tile_nl(P.edges, P.rowedge_idx + ea, n_cand, winding, xx0, yy, hard, s_aa[wid], reinterpret_cast<StagedEdge*>(stage), s_prefix, AA_VOLUME, lane, ok);
        if (!ok) *P.error_flag = 1;
        if ((hard >> lane) & 1u) op[(size_t)ent.x * 32 + lane] = (uint8_t)v;
        mine &= ~hard;
        __syncwarp();
      }
    }
    if ((mine >> lane) & 1u) op[(size_t)ent.x * 32 + lane] = (uint8_t)aa_opacity(mytot, AA_VOLUME);
  }
}

// ------------------------------------------------------------------------------------
// Row compositor of three-phase frames for FLAT scenes (every leaf a direct member of the scene list or of the
// background list, plain-filled paths and primitives): the front-to-back loop of render.ml:1268-1335 for one pixel
// row of one cell per warp, lane = pixel column.  Scan conversion (k_pre_scan) and antialiasing (k_pre_aa*) are
// done; here 32 list entries at a time are tested against the covered-so-far word `u` (lane = entry), and the entries
// that still show are composited in list order: vis = S & u, edge pixels dissolve the fill by their opacity byte
// (render.ml:1201-1204), PreTrans dissolves again (1295-1298), acc = over acc s, u' = u - opaque (1294, 1308).
// One block per cell (CELL_H warps), cells in heavy-first order; background cells were finished by k_prefill.
// ------------------------------------------------------------------------------------
constexpr int COMP_WARPS = 8;   // rows of a cell per block
__global__ void __launch_bounds__(COMP_WARPS * 32) k_comp_rows(WalkParams P, const int2* __restrict__ item_attr, const int4* __restrict__ order, int n_blocks) {
  constexpr int PARTS = CELL_H / COMP_WARPS;   // blocks per cell
  // Block b takes position b / PARTS of the heavy-first order (blocks are dispatched in index order, so the long
  // lists start first).  The order was flattened by k_comp_order when the cells were binned: one load gives the
  // cell, its list and its header flags — every warp fetches it for itself, no barrier; the grid is sized for
  // every cell, and positions beyond the queued cells (background cells were finished by k_prefill) leave at once.
  // A resident grid walks the positions (n_blocks = cells x PARTS of them, most of them background cells that leave at
  // once: one block per position would spend the launch on block dispatch).
  cudaGridDependencySynchronize();   // the opacities of the antialiasing kernel
  for (int bq = blockIdx.x; bq < n_blocks; bq += gridDim.x) {
  const int q = bq / PARTS;
  int4 oc;
  if (order) oc = order[q];
  else { oc = make_int4(q < P.n_cells ? q : -1, 0, 0, 0); if (oc.x >= 0) { const int2 rg0 = P.cell_rng[q]; oc.y = rg0.x; oc.z = rg0.y; oc.w = ((P.cell_head ? P.cell_head[q].y : 0) & 15) | ((q % P.fr.cntx) << 4) | ((q / P.fr.cntx) << 19); } }
  const int cell = oc.x;
  if (cell < 0) { if (order) break; continue; }   // (ordered: every later position is empty too)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lbit = 1u << lane;
  const int tile = P.fr.ctx0 + ((oc.w >> 4) & 0x7FFF), by = (oc.w >> 19) & 0xFFF;
  const int y = (P.cell_row0 + by) * CELL_H + (bq % PARTS) * COMP_WARPS + wid, row = y & (CELL_H - 1);
  if (y < P.fr.band_y0 || y >= P.fr.band_y1) continue;
  uint32_t colmask = interval_mask32(tile * TILE_W, P.ux0, P.ux1);
  if (tile * TILE_W + 31 >= P.fr.W) colmask &= interval_mask32(tile * TILE_W, 0, P.fr.W - 1);
  uint32_t u = P.u_init ? (P.u_init[(size_t)y * P.fr.tiles_x + tile] & colmask) : ((y >= P.uy0 && y <= P.uy1) ? colmask : 0u);
  const uint32_t u_update = u;
  uint32_t* u_rec = P.u_out ? P.u_out + (size_t)y * P.fr.tiles_x + tile : nullptr;   // receives u after the scene list
  if (u == 0u) { if (u_rec && lane == 0) *u_rec = 0u; continue; }
  if ((oc.w & 1) && !P.resume) {
    // a background cell that k_prefill did not take (frames mirrored to peer framebuffers spread these stores over
    // the compositor's blocks): one opaque primitive covers the cell, nothing was scan-converted for it
    if (u_rec && lane == 0) *u_rec = (oc.w & 2) ? 0u : u;
    if (P.touched && lane == 0) P.touched[(size_t)y * P.fr.tiles_x + tile] |= u;
    if (u & lbit) {
      const size_t at = (size_t)y * P.fr.W + tile * TILE_W + lane;
      const uint32_t bg = (uint32_t)P.cell_head[cell].x;
      P.fb[at] = bg;
      for (int k = 0; k < P.n_peers; k++) P.peer_fb[k][at] = bg;
    }
    continue;
  }
  const int2 rg = make_int2(oc.y, oc.z);
  const uint2* sc_row = P.pre_sc + row;
  const uint8_t* op_row = P.pre_op + (size_t)row * 32 + lane;
  uint32_t acc = 0u, touched_w = 0u;
  // a pass that continues a frame (after a filter, render.ml:1080-1131): the accumulator carries on from the framebuffer
  if (P.resume && (u & lbit)) acc = P.fb[(size_t)y * P.fr.W + tile * TILE_W + lane];
  for (int base = rg.x; base < rg.y; base += 32) {
    if (u == 0u) break;   // nothing of this row is uncovered any more (render.ml:1321-1322)
    const int it = base + lane;
    uint2 sc = make_uint2(0u, 0u); int2 at = make_int2(0, 0);
    if (it < rg.y) { sc = sc_row[(size_t)it * CELL_H]; at = item_attr[it]; }
    unsigned hits = __ballot_sync(0xFFFFFFFFu, (sc.x & u) != 0u);
    // the scene list ends where the background list begins: u is recorded between the two
    unsigned bgm = u_rec ? __ballot_sync(0xFFFFFFFFu, (at.y & 2) != 0) : 0u;
    unsigned later = bgm ? (hits & bgm) : 0u;
    hits &= ~later;
    for (;;) {
      while (hits) {
        const int k = __ffs((int)hits) - 1; hits &= hits - 1;
        const uint32_t S = __shfl_sync(0xFFFFFFFFu, sc.x, k);
        const uint32_t vis = S & u;
        if (vis == 0u) continue;
        touched_w |= vis;
        const uint32_t C = __shfl_sync(0xFFFFFFFFu, sc.y, k);
        const uint32_t c0 = (uint32_t)__shfl_sync(0xFFFFFFFFu, at.x, k); const int fl = __shfl_sync(0xFFFFFFFFu, at.y, k);
        // the common case — an opaque plain fill seen through its interior only (no edge pixel of this row still shows,
        // no dissolve): every visible pixel takes the colour under what is already there and is finished
        // (over a b with b opaque is opaque; over clear b = b, colour.ml:314-316) — no opacity bytes, no vote
        if ((c0 >> 24) == 255u && (fl & ~7) == 0 && !((fl & 1) && (vis & C))) {
          const bool under = (vis & lbit) && acc != 0u;
          if (__any_sync(0xFFFFFFFFu, under)) { if (vis & lbit) acc = px_over(acc, c0); }
          else if (vis & lbit) acc = c0;
          u &= ~vis;
          continue;
        }
        uint32_t col = c0;
        // shptorender ∩ maxshape = vis & ~(S & ~C) = vis & C: those pixels dissolve the fill by their opacity
        if ((fl & 1) && (vis & C & lbit)) col = px_dissolve(c0, op_row[(size_t)(base + k) * (CELL_H * 32)]);
        if (fl & 8) {
          // a Convolved object: the pre-convolved sprite, cropped to the visible max-shape (render.ml:1052)
          const ObjRec& o = P.objs[P.cell_items[base + k]];
          if (vis & C & lbit) col = P.conv_px[(size_t)o.cv_px + (size_t)(y - o.dy - o.cv_y0) * (o.cv_nw * 32) + (tile * TILE_W + lane - o.dx - o.cv_x0)];
        }
        if (fl >> 8) col = px_dissolve(col, (fl >> 8) - 1);
        if (vis & lbit) acc = px_over(acc, col);
        u &= ~__ballot_sync(0xFFFFFFFFu, (vis & lbit) && (acc >> 24) == 255u);
      }
      if (!bgm) break;
      if (lane == 0) *u_rec = u;
      u_rec = nullptr; bgm = 0u;
      hits = later;
    }
  }
  if (u_rec && lane == 0) *u_rec = u;   // the scene list ran out (or u did) before any member of the background list
  if (P.touched && touched_w && lane == 0) P.touched[(size_t)y * P.fr.tiles_x + tile] |= touched_w;
  if ((u_update & lbit) && (P.write_clear || acc != 0u)) {
    const size_t at = (size_t)y * P.fr.W + tile * TILE_W + lane;
    P.fb[at] = acc;
    for (int k = 0; k < P.n_peers; k++) P.peer_fb[k][at] = acc;
  }
  }
}

constexpr int SCAN_CHUNK_WORDS = 8;  // one thread scans a 256-pixel window of one row
// With `U` and `T` (planes laid out like S): T = S & U, the part of the shape inside an update set (render.ml:1281).
__global__ void k_scan_rows(const EdgeRec* __restrict__ edges, int n_edges, int winding, int y0, int n_rows,
                            int wx0, int nw, uint32_t* __restrict__ S, uint32_t* __restrict__ C, int* error_flag,
                            const uint32_t* __restrict__ U = nullptr, uint32_t* __restrict__ T = nullptr) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  int w0 = blockIdx.y * SCAN_CHUNK_WORDS;
  if (r >= n_rows || w0 >= nw) return;
  SinkMem sink;
  sink.wx0 = wx0 + 32 * w0; sink.nwords = min(SCAN_CHUNK_WORDS, nw - w0); sink.stride = 1;
  sink.S = S + (size_t)r * nw + w0; sink.C = C + (size_t)r * nw + w0;
  const int words = sink.nwords, x0 = sink.wx0;
  if (!scan_row(edges, nullptr, n_edges, 1, y0 + r, winding, false, sink.wx0, sink.wx0 + 32 * sink.nwords - 1, sink)) {
    // more crossings touch the 256-pixel window than a list holds (a line of small text as one compound path): word by word
    for (int w = 0; w < words; w++) {
      sink.wx0 = x0 + 32 * w; sink.nwords = 1;
      sink.S = S + (size_t)r * nw + w0 + w; sink.C = C + (size_t)r * nw + w0 + w;
      if (!scan_row(edges, nullptr, n_edges, 1, y0 + r, winding, false, sink.wx0, sink.wx0 + 31, sink)) *error_flag = 1;
    }
  }
  if (T) for (int w = 0; w < words; w++) { const size_t i = (size_t)r * nw + w0 + w; T[i] = S[i] & U[i]; }
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
