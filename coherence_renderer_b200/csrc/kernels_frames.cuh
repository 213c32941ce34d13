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
