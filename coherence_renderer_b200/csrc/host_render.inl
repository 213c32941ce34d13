// host_render.inl — part of coherence_b200.cu (one translation unit; included in order): render_pass (binning + walker launches), filter passes, coh_render_frame.

// One walk over the leaves [l0, l1) of a scene: binning + k_walk.
struct PassArgs {
  int l0, l1;                 // leaf range (list order)
  int ux, uy, uw, uh;         // update box (used when u_init is null)
  const uint32_t* u_init;     // update set as a bit-frame, or null
  uint32_t* u_out;            // receives `u` after the scene list, or null (may alias u_init)
  uint32_t* fb;               // target canvas
  bool write_clear, resume;
  bool collapsed = false;     // walk the leaf list in which cached objects are one sprite leaf each (l0 / l1 index that list)
};
static int render_pass(coh_ctx* ctx, DevScene* s, const PassArgs& A) {
  Frame fr = ctx->fr;
  // the band is the context's share of the FRAME: a pass into a canvas of its own (a filter's reading scene, which
  // reaches rows outside the band — halo rows are recomputed, not exchanged) covers whatever it is asked for
  if (A.fb != ctx->fb) { fr.band_y0 = 0; fr.band_y1 = fr.H; }
  const int ux = A.ux, uy = A.uy, uw = A.uw, uh = A.uh;
  const bool write_clear = A.write_clear;
  LeafView& V = A.collapsed ? s->sp : s->full;
  const int extras = A.collapsed ? std::max(s->extras, 1) : s->extras;   // sprite leaves are read like Convolved objects' canvases
  const int n_leaves = A.l1 - A.l0;
  const int4* leaf_box = V.leaf_box + A.l0;
  const int* leaves = V.leaves + A.l0;
  if (fr.band_y1 <= fr.band_y0 || uw <= 0 || uh <= 0) return 0;
  // only the cell rows the update box reaches (a dirty region is usually a small part of the frame)
  const int ry0 = std::max(fr.band_y0, uy), ry1 = std::min(fr.band_y1, uy + uh);
  if (ry1 <= ry0) return 0;
  // ... and only the tile columns it reaches
  fr.ctx0 = std::max(0, ux >> 5);
  const int ctx1 = std::min(fr.tiles_x - 1, (int)(((long long)ux + uw - 1) >> 5));
  if (ctx1 < fr.ctx0) return 0;
  fr.cntx = ctx1 - fr.ctx0 + 1;
  const bool whole = A.l0 == 0 && A.l1 == V.n && ry0 == fr.band_y0 && ry1 == fr.band_y1 && fr.cntx == fr.tiles_x;
  int cell_row0 = ry0 / CELL_H, cell_row1 = (ry1 - 1) / CELL_H;
  if (A.u_out && A.u_out != A.u_init && !(ry0 == fr.band_y0 && ry1 == fr.band_y1 && fr.cntx == fr.tiles_x))
    // rows and columns the walk does not visit have nothing uncovered
    CK(cudaMemsetAsync(A.u_out + (size_t)fr.band_y0 * fr.tiles_x, 0, 4 * (size_t)(fr.band_y1 - fr.band_y0) * fr.tiles_x, ctx->stream));
  int n_cells = (cell_row1 - cell_row0 + 1) * fr.cntx;
  const bool big = n_leaves > 1024;
  const bool ordered = !s->has_fancy;  // with fancy fills the queue must stay row-major (carry look-back)
  if (!ctx->queue) {
    CK(DMALLOC(&ctx->queue, sizeof(int)));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->n_sms = prop.multiProcessorCount;
  }
  // capacity of the item pool: the exact total is a pure function of the object boxes and the
  // frame geometry, so it is computed on the host (once per scene and geometry) — no device
  // round trip inside a frame.
  size_t total = 0;
  if (!whole || V.items_for_W != fr.W || V.items_for_H != fr.H || V.items_for_y0 != fr.band_y0 || V.items_for_y1 != fr.band_y1) {
    size_t tot = 0;
    for (int li = A.l0; li < A.l1; li++) {
      const ObjRec& o = s->h_objs[V.h_leaves[li]];
      int cx0 = std::max(o.bx0 >> 5, fr.ctx0), cx1 = std::min(o.bx1 >> 5, ctx1);
      int cy0 = std::max(floordiv(o.by0, CELL_H), cell_row0), cy1 = std::min(floordiv(o.by1, CELL_H), cell_row1);
      if (cx1 >= cx0 && cy1 >= cy0) tot += (size_t)(cx1 - cx0 + 1) * (cy1 - cy0 + 1);
    }
    total = tot;
    if (whole) { V.coarse_total_valid = false; V.items_total = tot; V.items_for_W = fr.W; V.items_for_H = fr.H; V.items_for_y0 = fr.band_y0; V.items_for_y1 = fr.band_y1; }
  } else total = V.items_total;
  const size_t need = total;
  // persistent grid: exactly one resident wave.  Work items are 4 rows high, or 16 for very large scenes.
  int walk_h = big ? 16 : 4;
  // Few cells (a band of an 8-GPU split, a small dirty region): the launch is bounded by its longest work item,
  // not by throughput — one-row items shorten that path (measured on the lion at 8 GPUs: 0.164 -> 0.141 ms;
  // at 1 to 4 GPUs four-row items are as fast or faster).
  if (!big && (long long)n_cells * 4 < 3LL * ctx->n_sms * WALK_MIN_CTAS * WALK_WARPS) walk_h = 1;
  // ... and a band of a very large scene with fewer than three 16-row items per resident warp is bounded by its heaviest
  // cells: four-row items cost more work in total (C3 whole frame: 4.3 vs 3.0 ms) but quarter the critical path
  // (C3, slowest 1/8 band: 0.98 -> 0.73 ms; at 1/4 the two are equal)
  if (big && (long long)n_cells < 3LL * ctx->n_sms * WALK_MIN_CTAS * WALK_WARPS) walk_h = 4;
  if (ctx->opt_walk_h) walk_h = ctx->opt_walk_h;  // tests force every variant
  // Plain-filled paths and primitives only, a list pool of moderate size: three-phase frame (kernels.cuh)
  // (small launches — a band of an 8-GPU split, a dirty region — stay fused: four dependent launches cost more
  // than the parallelism gains there; measured on 1/8 bands of the lion: 0.051 vs 0.059 ms)
  const int force = ctx->opt_fused;   // tests force either path: 1 fused, 0 three-phase
  // eligible: every leaf of the range is a path, a primitive or a Convolved object
  bool kinds_ok = extras == 0;
  bool has_conv = false;
  if (!kinds_ok && !big) {
    kinds_ok = true;
    for (int li = A.l0; li < A.l1 && kinds_ok; li++) {
      const int k = s->h_objs[V.h_leaves[li]].kind;
      kinds_ok = k == K_PATH || k == K_PRIM || k == K_CONV;
      has_conv = has_conv || k == K_CONV;
    }
  }
  // (a pass that continues a frame — after a filter — takes the three-phase path too: its compositing walk is the
  // variant that starts the root accumulators from the framebuffer)
  const bool pre = kinds_ok && !big && total > 0 && total * CELL_H <= (size_t)(1 << 23) &&
                   // flat scenes have the row compositor: worth it from a few thousand pairs on (measured on bands of the
                   // lion: 1/8 of the frame 0.069 -> 0.053 ms, 1/2 0.155 -> 0.098 ms); other scenes composite with the
                   // walker, and their small passes (a drag's dirty region: 0.088 vs 0.103 ms) stay fused
                   (force >= 0 ? force == 0 : ((s->flat_ok && !A.collapsed) ? (long long)total * CELL_H >= ctx->opt_pre_min_pairs : (walk_h != 1 || (long long)total * CELL_H >= ctx->opt_pre_min_pairs_walk)));
  // Background cells of a box update are finished by k_prefill and never enter the walker's queue.  With peer
  // framebuffers only in three-phase frames, where the prefill — and its mirrored stores over NVLink — runs beside the
  // scan kernels (in a fused walk the mirrored stores of background cells are better spread over the walker's warps:
  // measured at 2 / 4 / 8 GPUs)
  const bool mirrored = A.fb == ctx->fb && ctx->n_peers > 0;
  const bool prefill = !big && ordered && !A.u_init && !A.resume && (!mirrored || (pre && ctx->opt_fork_prefill && !(ctx->opt_ab & 4)));
  // Whole-frame binning of a small scene is kept with the scene; every other pass bins into the context's scratch.
  const bool keep = whole && !big && ctx->opt_bin_cache;
  BinSet& B = keep ? V.bins : ctx->bins;
  const int key[6] = {fr.W, fr.H, fr.band_y0, fr.band_y1, ordered ? 1 : 0, prefill ? 1 : 0};
  const bool hit = keep && V.bins_valid && memcmp(key, V.bins_key, sizeof key) == 0;
  if (n_cells > B.n_cells_cap) {
    DFREE(B.cell_order); DFREE(B.cell_head); DFREE(B.cell_rng); DFREE(B.comp_order);
    B.n_cells_cap = 0;
    CK(DMALLOC(&B.cell_head, sizeof(int2) * n_cells));
    CK(DMALLOC(&B.cell_rng, sizeof(int2) * n_cells));
    CK(DMALLOC(&B.cell_order, sizeof(int) * (size_t)n_cells * BIN_CLASSES));  // one-pass binning keeps one segment per length class
    CK(DMALLOC(&B.comp_order, sizeof(int4) * (size_t)n_cells));
    B.n_cells_cap = n_cells;
  }
  if (!B.state) CK(DMALLOC(&B.state, sizeof(int) * ORDER_BINS));
  CK(cudaMemsetAsync(ctx->queue, 0, sizeof(int), ctx->stream));
  if (ctx->timing) { if (drain_timing(ctx)) return 1; CK(cudaEventRecord(ctx->ev[0], ctx->stream)); }
  if (need > B.cell_items_cap) {
    DFREE(B.cell_items); DFREE(B.item_cell); DFREE(B.item_attr); DFREE(B.item_rec);
    B.cell_items_cap = 0;
    size_t cap = need + need / 2 + 1024;
    CK(DMALLOC(&B.cell_items, sizeof(int) * cap));
    CK(DMALLOC(&B.item_cell, sizeof(int) * cap));
    CK(DMALLOC(&B.item_attr, sizeof(int2) * cap));
    CK(DMALLOC(&B.item_rec, sizeof(int4) * 2 * cap));
    B.cell_items_cap = cap;
  }
  // K1.  Small scenes: warp per cell scanning all leaves (lists come out sorted, no atomics).  Large scenes:
  // warp per leaf over the coarse cells it covers, then every fine cell from its coarse list.
  if (hit) {
    // nothing to do: lists, classes and cell headers of this geometry are resident
  } else if (!big) {
    CK(cudaMemsetAsync(B.state, 0, sizeof(int) * ORDER_BINS, ctx->stream));
    // one pass: hit masks in registers, lists carved from one cursor, length classes instead of a sort
    const int bin_blocks = cdiv(n_cells * 32, 256);
    k_bin1<<<bin_blocks, 256, 0, ctx->stream>>>(leaf_box, leaves, n_leaves, fr, cell_row0, n_cells, B.cell_rng, B.cell_items, B.state,
                                               ordered ? B.cell_order : nullptr, s->objs, B.cell_head, B.item_cell, prefill ? 1 : 0, s->attr, B.item_attr, B.item_rec); LAUNCHED();
    B.comp_valid = false;
    if (keep) { V.bins_valid = true; memcpy(V.bins_key, key, sizeof key); }
  } else {
    CK(cudaMemsetAsync(B.state, 0, sizeof(int) * ORDER_BINS, ctx->stream));
    // two levels: leaves into coarse cells (object-parallel, sorted per coarse list), then every fine cell from its coarse list
    const int ctx_x = cdiv(fr.tiles_x, COARSE), crow0 = cell_row0 >> COARSE_SHIFT, crow1 = cell_row1 >> COARSE_SHIFT;
    const int n_coarse = ctx_x * (crow1 - crow0 + 1);
    size_t ctot = 0;
    if (whole && V.coarse_total_valid) ctot = V.coarse_total;   // a pure function of the boxes and the frame geometry, like items_total
    else {
      for (int li = A.l0; li < A.l1; li++) {
        const ObjRec& o = s->h_objs[V.h_leaves[li]];
        int cx0 = std::max(floordiv(o.bx0, 32 * COARSE), 0), cx1 = std::min(floordiv(o.bx1, 32 * COARSE), ctx_x - 1);
        int cy0 = std::max(floordiv(o.by0, CELL_H * COARSE), crow0), cy1 = std::min(floordiv(o.by1, CELL_H * COARSE), crow1);
        if (cx1 >= cx0 && cy1 >= cy0) ctot += (size_t)(cx1 - cx0 + 1) * (cy1 - cy0 + 1);
      }
      if (whole) { V.coarse_total = ctot; V.coarse_total_valid = true; }
    }
    if (2 * ctot + 1 > ctx->coarse_cap || (size_t)n_coarse + 1 > ctx->coarse_cells_cap) {
      DFREE(ctx->coarse_items); DFREE(ctx->coarse_counts); DFREE(ctx->coarse_off);
      ctx->coarse_cap = 0; ctx->coarse_cells_cap = 0;
      const size_t cap = 2 * ctot + ctot / 2 + 1024, ccap = (size_t)n_coarse + 1;
      CK(DMALLOC(&ctx->coarse_items, sizeof(int) * cap));
      CK(DMALLOC(&ctx->coarse_counts, sizeof(int) * ccap));
      CK(DMALLOC(&ctx->coarse_off, sizeof(int) * (ccap + 1)));
      ctx->coarse_cap = cap; ctx->coarse_cells_cap = ccap;
    }
    const int obj_blocks = cdiv(std::max(n_leaves, 1) * 32, 256);
    CK(cudaMemsetAsync(ctx->coarse_counts, 0, sizeof(int) * n_coarse, ctx->stream));
    k_bin_obj<false><<<obj_blocks, 256, 0, ctx->stream>>>(leaf_box, n_leaves, ctx_x, crow0, crow1, ctx->coarse_counts, nullptr, nullptr); LAUNCHED();
    if (exclusive_scan(ctx, ctx->coarse_counts, ctx->coarse_off, n_coarse, nullptr)) return 1;
    CK(cudaMemsetAsync(ctx->coarse_counts, 0, sizeof(int) * n_coarse, ctx->stream));
    k_bin_obj<true><<<obj_blocks, 256, 0, ctx->stream>>>(leaf_box, n_leaves, ctx_x, crow0, crow1, ctx->coarse_counts, ctx->coarse_off, ctx->coarse_items); LAUNCHED();
    k_bin_sort<<<cdiv(n_coarse * 32, 128), 128, 0, ctx->stream>>>(ctx->coarse_off, ctx->coarse_items, ctx->coarse_items + ctot, n_coarse); LAUNCHED();
    k_bin2<<<cdiv(n_cells * 32, 256), 256, 0, ctx->stream>>>(leaf_box, leaves, ctx->coarse_off, ctx->coarse_items, ctx_x, crow0, fr, cell_row0, n_cells, B.cell_rng,
                                                          B.cell_items, B.state, ordered ? B.cell_order : nullptr); LAUNCHED();
  }
  WalkParams P;
  P.objs = s->objs; P.edges = s->edges; P.points = s->points; P.stamps = s->stamps;
  P.rowedge_ptr = s->rowedge_ptr; P.rowedge_idx = s->rowedge_idx; P.brush_ranges = s->brush_ranges;
  P.conv_bits = s->conv_bits; P.conv_px = s->conv_px;
  P.cell_rng = B.cell_rng;
  P.cls_cells = ordered ? B.cell_order : nullptr; P.cls_cnt = ordered ? B.state + 1 : nullptr;
  P.cell_items = B.cell_items; P.cell_head = big ? nullptr : B.cell_head; P.aa = ctx->d_aa; P.fr = fr; P.cell_row0 = cell_row0;
  P.ux0 = ux; P.uy0 = uy; P.ux1 = ux + uw - 1; P.uy1 = uy + uh - 1;
  P.u_init = A.u_init; P.u_out = A.u_out; P.fb = A.fb; P.error_flag = ctx->d_error;
  P.write_clear = write_clear ? 1 : 0; P.resume = A.resume ? 1 : 0;
  P.touched = ctx->touched;
  P.n_peers = (A.fb == ctx->fb) ? ctx->n_peers : 0;   // only the frame itself is mirrored, not filter canvases
  for (int k = 0; k < COH_MAX_PEERS; k++) P.peer_fb[k] = k < P.n_peers ? ctx->peer_fb[k] : nullptr;
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
    if (extras == 0) LAUNCH_WALK_E(CARRYV, 0); else if (extras == 1) LAUNCH_WALK_E(CARRYV, 1); else LAUNCH_WALK_E(CARRYV, 2); \
  } while (0)
  P.queue = ctx->queue; P.n_cells = n_cells;
  P.pre_sc = nullptr; P.pre_op = nullptr; P.item_cell = nullptr;
  if (ctx->timing) CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  // background cells: cleared to the background colour by their own kernel.  In a three-phase frame it touches nothing
  // the scan kernels read, so it runs beside them on a second stream and joins before the compositor.
  bool prefill_forked = false;
  if (prefill) {
    BinPrefill pf; memset(&pf, 0, sizeof pf);
    pf.fb = A.fb; pf.u_out = A.u_out; pf.ux0 = ux; pf.uy0 = uy; pf.ux1 = ux + uw - 1; pf.uy1 = uy + uh - 1;
    pf.n_peers = mirrored ? ctx->n_peers : 0;
    for (int k = 0; k < pf.n_peers; k++) pf.peer_fb[k] = ctx->peer_fb[k];
    cudaStream_t st = ctx->stream;
    if (pre && ctx->opt_fork_prefill) {
      if (!ctx->aux_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
      }
      CK(cudaEventRecord(ctx->ev_fork, ctx->stream)); CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
      st = ctx->aux_stream; prefill_forked = true;
    }
    k_prefill<<<cdiv(n_cells * 32, 256), 256, 0, st>>>(B.cell_head, fr, cell_row0, n_cells, pf); LAUNCHED();
    if (prefill_forked) CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
  }
  if (pre) {
    const size_t n_pairs = total * CELL_H;
    if (n_pairs > ctx->pre_cap) {
      DFREE(ctx->pre_sc); DFREE(ctx->pre_list); DFREE(ctx->pre_op);
      const size_t cap = n_pairs + n_pairs / 4 + 1024;
      CK(DMALLOC(&ctx->pre_sc, sizeof(uint2) * cap));
      CK(DMALLOC(&ctx->pre_list, sizeof(int4) * cap)); CK(DMALLOC(&ctx->pre_op, 32 * cap));
      if (!ctx->pre_n) CK(DMALLOC(&ctx->pre_n, 2 * sizeof(int)));   // list length, next pair of the antialiasing kernel
      ctx->pre_cap = cap;
    }
    P.item_cell = B.item_cell;
    // The four kernels of the chain are launched with programmatic stream serialization (PDL): every producer signals at
    // its start that its dependents may be scheduled, every consumer does what does not depend on the producer (staging
    // the AA table, its own indices) and waits in cudaGridDependencySynchronize () for the producer's grid to complete —
    // the consumer's blocks fill the SMs as the producer's last blocks leave, instead of after the grid has drained.
    const bool pdl = ctx->opt_pdl;
    k_pre_scan<<<cdiv((int)n_pairs, 128), 128, 0, ctx->stream>>>(P, (int)n_pairs, ctx->pre_sc, B.item_rec, ctx->pre_n); LAUNCHED();
    CK(launch_chain(k_pre_vis, cdiv(n_cells * CELL_H, 128), 128, ctx->stream, pdl && !(ctx->opt_ab & 64), P, (const uint2*)ctx->pre_sc, ctx->pre_list, ctx->pre_n, (const int2*)B.item_attr)); LAUNCHED();
    if (ctx->aa_general) { k_pre_aa<<<ctx->n_sms * 4, 256, 0, ctx->stream>>>(P, ctx->pre_list, ctx->pre_n, ctx->pre_op); LAUNCHED(); }
    else {
      // interval form (the few pairs with rows that are not runs take the bit-row routine inside the same kernel)
      CK(launch_chain(k_pre_aa_runs, ctx->n_sms * 4, AA2_WARPS * 32, ctx->stream, pdl && !(ctx->opt_ab & 32), P, (const int4*)ctx->pre_list, (const int*)ctx->pre_n, ctx->pre_op, ctx->pre_n + 1)); LAUNCHED();
    }
    P.pre_sc = ctx->pre_sc; P.pre_op = ctx->pre_op;
    // the row compositor touches no cell that k_prefill finishes (and the other way round): with PDL the join moves behind it,
    // so that no event wait stands between the antialiasing kernel and its dependent
    const bool pdl_comp = pdl && !(ctx->opt_ab & 16);
    const bool join_late = prefill_forked && pdl_comp && !(ctx->opt_ab & 8) && !ctx->aa_general && s->flat_ok && !A.collapsed && ctx->opt_comp_rows;
    if (prefill_forked && !join_late) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    const int pgrid = std::min(ctx->n_sms * WALK_MIN_CTAS, cdiv(n_cells * (CELL_H / 4), WALK_WARPS));
    if (s->flat_ok && !A.collapsed && ctx->opt_comp_rows) {
      // flat scene: one warp per pixel row of a cell composites the pre-scanned, pre-antialiased list entries
      // (the heavy-first order flattened for this kernel: made when the cells were binned, kept with cached lists)
      // (a small pass — a filter's lens — gains nothing from the heavy-first order: one dependent launch less)
      const bool comp_ordered = ordered && n_cells > 2048;
      if (comp_ordered && !B.comp_valid) { k_comp_order<<<cdiv(n_cells, 256), 256, 0, ctx->stream>>>(B.state + 1, B.cell_order, B.cell_rng, B.cell_head, n_cells, B.comp_order, fr.cntx); LAUNCHED(); B.comp_valid = true; }
      const int comp_blocks = n_cells * (CELL_H / COMP_WARPS);
      CK(launch_chain(k_comp_rows, std::min(comp_blocks, ctx->n_sms * (2048 / (COMP_WARPS * 32))), COMP_WARPS * 32, ctx->stream, pdl_comp && (join_late || !prefill_forked), P, (const int2*)B.item_attr, (const int4*)(comp_ordered ? B.comp_order : nullptr), comp_blocks)); LAUNCHED();
      if (join_late) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    } else if (s->has_fancy) {  // fancy fills: the compositing walk keeps the cross-tile carry (row-major queue order)
      // (the touched plane of a smear filter is compiled into the EXTRAS = 2 variants only)
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
      if (A.resume || ctx->touched) k_walk<true, 2, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P);
      else if (has_conv) k_walk<true, 1, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P);
      else k_walk<true, 0, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P);
      LAUNCHED();
    } else {
      if (A.resume || ctx->touched) k_walk<false, 2, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P);
      else if (has_conv) k_walk<false, 1, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P);
      else k_walk<false, 0, 4, true><<<pgrid, WALK_WARPS * 32, 0, ctx->stream>>>(P);
      LAUNCHED();
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
// Frames of scenes with cached sprites (render.ml:1169-1242 spriteof; cache.ml:328-367, 390-407).  For every cached
// object: shptorender = r' - pshape — here update ∩ shape - pshape, a superset of the reference's (it intersects with
// the exact u; rendering more pixels of a plain-filled group into the cache changes no value) — is rendered from
// the object's members alone into a scratch canvas and merged into the object's canvas and pshape
// (Cache.addsprite); then the frame is walked over the list in which the object is one sprite leaf.  Once pshape has
// reached the shape (one asynchronous counter read-back) the object costs one leaf per frame, wherever it is moved.
// ---------------------------------------------------------------------------------------
static bool sprites_usable(coh_ctx* ctx, DevScene* s) {
  if (!ctx->usecache || s->sprites.empty() || !s->filters.empty()) return false;
  if (ctx->fr.band_y0 != 0 || ctx->fr.band_y1 != ctx->fr.H) return false;   // a band never completes a sprite that straddles it
  for (const SpriteEntry& e : s->sprites) if (e.dead) return false;
  return true;
}
static int render_frame_passes(coh_ctx* ctx, DevScene* s, PassArgs A) {
  if (!sprites_usable(ctx, s)) return render_pass(ctx, s, A);
  const Frame& fr = ctx->fr;
  const int nw = fr.tiles_x;
  bool filled = false;
  for (SpriteEntry& e : s->sprites) {
    if (e.ev_pending && cudaEventQuery(e.ev) == cudaSuccess) { e.ev_pending = false; if (*e.h_missing == 0) e.complete = true; }
    if (e.complete) continue;
    const ObjRec& o = s->h_objs[e.leaf_rec];
    const int x0 = std::max(std::max(A.ux, o.bx0), 0), x1 = std::min(std::min(A.ux + A.uw - 1, o.bx1), fr.W - 1);
    const int y0 = std::max(std::max(A.uy, o.by0), 0), y1 = std::min(std::min(A.uy + A.uh - 1, o.by1), fr.H - 1);
    if (x1 < x0 || y1 < y0) continue;
    const int rows = y1 - y0 + 1;
    uint32_t *T = nullptr, *tmp = nullptr;
    CK(DMALLOC(&T, 4 * (size_t)nw * fr.H));
    CK(DMALLOC(&tmp, 4 * (size_t)fr.W * fr.H));
    const uint32_t* S = s->conv_bits + o.cv_bits;
    k_sprite_todo<<<dim3(cdiv(nw, 128), rows), 128, 0, ctx->stream>>>(A.u_init, A.ux, A.uy, A.ux + A.uw - 1, A.uy + A.uh - 1, S, e.valid, o.cv_x0, o.cv_y0, o.cv_nw, o.cv_h,
                                                                      o.dx, o.dy, fr.W, nw, y0, rows, T, nullptr); LAUNCHED();
    PassArgs F{e.l0, e.l1, x0, y0, x1 - x0 + 1, rows, T, nullptr, tmp, true, false};
    if (render_pass(ctx, s, F)) return 1;
    k_sprite_store<<<dim3(cdiv(fr.W, 256), rows), 256, 0, ctx->stream>>>(tmp, T, s->conv_px + o.cv_px, e.valid, o.cv_x0, o.cv_y0, o.cv_nw, o.cv_h, o.dx, o.dy, fr.W, nw, y0, rows); LAUNCHED();
    if (!e.ev_pending) {   // has pshape reached the shape?  (answered a few frames later, without a wait)
      CK(cudaMemsetAsync(e.d_missing, 0, sizeof(int), ctx->stream));
      k_sprite_missing<<<(unsigned)((e.plane_words + 255) / 256), 256, 0, ctx->stream>>>(S, e.valid, e.plane_words, e.d_missing); LAUNCHED();
      CK(cudaMemcpyAsync(e.h_missing, e.d_missing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaEventRecord(e.ev, ctx->stream));
      e.ev_pending = true;
    }
    DFREE(T); DFREE(tmp);
    filled = true;
  }
  if (filled) s->sprite_fills++; else s->sprite_hits++;
  A.collapsed = true; A.l0 = 0; A.l1 = s->sp.n;
  return render_pass(ctx, s, A);
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
static int object_shape_rec(coh_ctx* ctx, DevScene* s, int r, coh_shape_t* shape, coh_shape_t* minshape);
static int subscene_shape(coh_ctx* ctx, DevScene* ss, coh_shape_t* out);
static int subscene_render(coh_ctx* ctx, DevScene* ss, int nw, int h, uint32_t* A);
// temporaries of one filter application: released on every way out
struct StreamTemps {
  coh_ctx* ctx; std::vector<void*> v;
  explicit StreamTemps(coh_ctx* c) : ctx(c) {}
  ~StreamTemps() { for (void* p : v) cudaFreeAsync(p, ctx->stream); }
  cudaError_t get_(void** p, size_t bytes) { cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream); if (e == cudaSuccess) v.push_back(*p); return e; }
#define TMPGET(tmp, p, bytes) (tmp).get_((void**)(p), (bytes))
};

// `box` bounds the set bits of U: all work is confined to its rows (bit-frames are small; the RGBA8 canvases are only
// touched in the rows and columns the filter reads or writes).  `fresh`: nothing but filters has been composited into
// `target` so far — every pixel still in U has a clear accumulator there (a filter finishes the whole of its shape).
// Shape / coverage bit-rows of a filter's geometry and the antialiased opacity of all its pixels: computed at the first
// frame that needs them and kept with the scene (per frame geometry).
static int filter_geometry(coh_ctx* ctx, DevScene* s, DevScene::FilterRec& F) {
  const Frame& fr = ctx->fr;
  const int W = fr.W, H = fr.H, nw = fr.tiles_x;
  if (F.SG && F.gW == W && F.gH == H && F.gdx == F.dx && F.gdy == F.dy) return 0;
  DFREE(F.SG); DFREE(F.CG); DFREE(F.op);
  F.gW = W; F.gH = H; F.gdx = F.dx; F.gdy = F.dy;
  F.gy0 = std::max(F.by0, 0);
  F.gh = std::min(F.by1, H - 1) - F.gy0 + 1;
  if (F.gh <= 0) { F.gh = 0; return 0; }
  const int h = F.gh;
  const size_t nwords = (size_t)nw * h;
  CK(DMALLOC(&F.SG, 4 * nwords)); CK(DMALLOC(&F.CG, 4 * nwords)); CK(DMALLOC(&F.op, (size_t)nw * 32 * h));
  CK(cudaMemsetAsync(F.SG, 0, 4 * nwords, ctx->stream)); CK(cudaMemsetAsync(F.CG, 0, 4 * nwords, ctx->stream));
  if (F.geom_sub) {
    // COH_GEOM_NEXT: shape = the geometry object's (render.ml:472-474), matte = the alpha of its sprite (render.ml:1099)
    // — the object rendered as a scene of its own into a canvas (its members were aliased into canvas coordinates)
    DevScene* ss = F.geom_sub;
    coh_shape_t gs = 0;
    if (subscene_shape(ctx, ss, &gs)) return 1;
    if (gs) {
      const DevShape* G = (const DevShape*)gs;
      k_spans_to_bits<<<cdiv(h, 128), 128, 0, ctx->stream>>>(G->row_ptr, G->spans, G->y0 + F.gcy0 + F.dy, G->n_rows, F.gy0, h, -(F.gcx0 + F.dx), nw, F.SG); LAUNCHED();
      coh_shape_free(ctx, gs);
    }
    uint32_t* A = nullptr;
    const size_t cpx = (size_t)F.gcnw * 32 * F.gch;
    CK(DMALLOC(&A, 4 * cpx));
    CK(cudaMemsetAsync(A, 0, 4 * cpx, ctx->stream));
    if (subscene_render(ctx, ss, F.gcnw, F.gch, A)) return 1;
    k_canvas_alpha<<<dim3(cdiv(nw * 32, 256), h), 256, 0, ctx->stream>>>(A, F.gcnw * 32, F.gch, F.gcx0 + F.dx, F.gcy0 + F.dy, F.gy0, h, nw * 32, F.op); LAUNCHED();
    DFREE(A);
    return 0;
  }
  if (F.kind == COH_FILTER_SMEAR) {
    // geometry = the stroke's dummy brush (filters.ml:205-207): the boxes around its stamp points, opaque white all over
    const int side = 2 * F.brush_r + 1;
    k_stamp_boxes_to_bits<<<cdiv(F.count * side, 256), 256, 0, ctx->stream>>>(s->points + F.first, F.count, F.brush_r, F.gy0 - F.dy, h, -F.dx, nw, F.SG); LAUNCHED();
    CK(cudaMemsetAsync(F.op, 255, (size_t)nw * 32 * h, ctx->stream));
    return 0;
  }
  const EdgeRec* ed = s->edges + F.first;
  // shape of the geometry (render.ml:472-474) and its coverage (minshape = shape - coverage, needed for the matte)
  k_scan_rows<<<dim3(cdiv(h, 64), cdiv(nw, SCAN_CHUNK_WORDS)), 64, 0, ctx->stream>>>(ed, F.count, F.winding, F.gy0 - F.dy, h, -F.dx, nw, F.SG, F.CG, ctx->d_error); LAUNCHED();   // (an alias reads the geometry's own frame moved by whole pixels)
  // The geometry's matte (render.ml:1099-1103).  Polygon.polygon_sprite samples every pixel it is given, but a pixel
  // whose 5 x 5 neighbourhood lies in the geometry's minshape has no edge anywhere near its 2 x 2-pixel sampling window
  // (a minshape pixel's row band [32y-47, 32y+16] and its columns are free of edge pieces), so all 32 x 32 samples
  // are inside and the opacity is 255: only the rest is super-sampled (interior = erode 2 2 minshape, clipped 2 pixels
  // inside the rows / columns scanned here).
  uint32_t* Q = nullptr;
  CK(DMALLOC(&Q, 4 * nwords));
  if (F.aa_winding == F.winding) { k_matte_todo<<<dim3(cdiv(nw, 128), h), 128, 0, ctx->stream>>>(F.SG, F.CG, F.SG, Q, h, nw, W); LAUNCHED(); }
  else CK(cudaMemcpyAsync(Q, F.SG, 4 * nwords, cudaMemcpyDeviceToDevice, ctx->stream));   // a stroked path's sprite (EvenOdd) is not its shape's interior (NonZero): sample everything
  CK(cudaMemsetAsync(F.op, 255, (size_t)nw * 32 * h, ctx->stream));
  k_aa_rows<<<dim3(cdiv(nw, 8), h), 256, 0, ctx->stream>>>(ed, F.count, F.aa_winding, Q, F.gy0 - F.dy, h, -F.dx, nw, ctx->d_aa, F.op, ctx->d_error); LAUNCHED();
  DFREE(Q);
  return 0;
}
static int apply_filter(coh_ctx* ctx, DevScene* s, int fi, uint32_t* U, uint32_t* target, PixBox box, bool fresh) {
  DevScene::FilterRec& F = s->filters[fi];
  const Frame& fr = ctx->fr;
  const int W = fr.W, H = fr.H, nw = fr.tiles_x;
  const size_t nwords = (size_t)nw * H;
  // rows / columns of shptorender = shape(geometry) ∩ u
  const int y0 = std::max(std::max(F.by0, 0), box.y0), y1 = std::min(std::min(F.by1, H - 1), box.y1);
  const int x0 = std::max(std::max(F.bx0, 0), box.x0), x1 = std::min(std::min(F.bx1, W - 1), box.x1);
  if (y0 > y1 || x0 > x1) return 0;  // the geometry cannot meet u: nothing to render, nothing leaves u
  if (filter_geometry(ctx, s, F)) return 1;
  const int h = y1 - y0 + 1;
  const int m = F.kind == COH_FILTER_BLUR ? 2 * F.r + 1 : (F.kind == COH_FILTER_SMEAR ? F.brush_r : 0);   // reach of the reading shape
  const int ry0 = std::max(0, y0 - m), ry1 = std::min(H - 1, y1 + m), rh = ry1 - ry0 + 1;
  const PixBox tbox{x0, y0, x1, y1}, rbox{std::max(0, x0 - m), ry0, std::min(W - 1, x1 + m), ry1};
  const size_t po = (size_t)ry0 * W, pn = (size_t)rh * W;                 // canvas rows [ry0, ry1]
  const size_t r0 = (size_t)y0 * nw;                                      // first word of row y0 in a bit-frame
  const size_t g0 = (size_t)(y0 - F.gy0) * nw;                            // ... and in the geometry's own planes
  const uint32_t* SG = F.SG + g0;
  StreamTemps tmp(ctx);
  // bit-frames: T = shptorender, R = the reading shape, then the pixels the scene below shows through
  uint32_t* planes = nullptr;
  CK(TMPGET(tmp, &planes, 4 * nwords * 2));
  uint32_t *T = planes, *R = planes + nwords;
  CK(cudaMemsetAsync(planes, 0, 4 * nwords * 2, ctx->stream));
  // shptorender = r &&& u (render.ml:1281); a filter other than blur reads where it writes
  k_and_rows<<<(unsigned)(((size_t)h * nw + 255) / 256), 256, 0, ctx->stream>>>(SG, U + r0, T + r0, m > 0 ? nullptr : R + r0, (size_t)h * nw); LAUNCHED();
  if (F.kind == COH_FILTER_MINUS) {
    // filters.ml:291-303: the filter reads, and acts, only in shape (filter) ∩ shape (hd scene) ∩ shp
    const int hr = s->rec_of_abi[F.head_abi];
    coh_shape_t hs = 0, hm = 0;
    if (hr >= 0 && object_shape_rec(ctx, s, hr, &hs, &hm)) return 1;
    const DevShape* Hs = (const DevShape*)hs;
    if (Hs) {
      uint32_t* Hb = nullptr;
      CK(TMPGET(tmp, &Hb, 4 * (size_t)h * nw));
      CK(cudaMemsetAsync(Hb, 0, 4 * (size_t)h * nw, ctx->stream));
      k_spans_to_bits<<<cdiv(h, 128), 128, 0, ctx->stream>>>(Hs->row_ptr, Hs->spans, Hs->y0, Hs->n_rows, y0, h, 0, nw, Hb); LAUNCHED();
      k_bitop<<<(unsigned)(((size_t)h * nw + 255) / 256), 256, 0, ctx->stream>>>(T + r0, Hb, T + r0, (size_t)h * nw, 2); LAUNCHED();
    } else CK(cudaMemsetAsync(T + r0, 0, 4 * (size_t)h * nw, ctx->stream));
    CK(cudaMemcpyAsync(R + r0, T + r0, 4 * (size_t)h * nw, cudaMemcpyDeviceToDevice, ctx->stream));
    coh_shape_free(ctx, hs); coh_shape_free(ctx, hm);
  }
  if (ctx->touched) {   // an enclosing smear filter wants the shape of what is rendered here: this filter's sprite covers T
    k_bitop<<<(unsigned)(((size_t)h * nw + 255) / 256), 256, 0, ctx->stream>>>(ctx->touched + r0, T + r0, ctx->touched + r0, (size_t)h * nw, 0); LAUNCHED();
  }
  // The scene below renders the same pixels whatever region it is asked for (plain fills: no span-start quirk,
  // polygon.ml:736), so where the reading scene IS the scene below (monochrome, blur: filters.ml:229-258) the pixels
  // that show through the matte (render.ml:1105-1110) are taken from the reading scene's render before its filter function.
  const bool z_is_x = (F.kind == COH_FILTER_MONOCHROME || F.kind == COH_FILTER_BLUR) && !s->has_fancy;
  // reading scene -> X -> filter function -> Y
  uint32_t *X = nullptr, *Y = nullptr, *Z = nullptr;
  int blend_flags = fresh ? 1 : 0;
  if (F.kind != COH_FILTER_HOLE) {
    CK(TMPGET(tmp, &X, 4 * (size_t)W * H));
    CK(cudaMemset2DAsync(X + po + rbox.x0, 4 * (size_t)W, 0, 4 * (size_t)(rbox.x1 - rbox.x0 + 1), rh, ctx->stream));
    if (m > 0) {  // blur: read in bloat (2r+1) (2r+1) shp (filters.ml:247-250); smear: bloat rx ry shp (filters.ml:209) (T is empty outside [y0, y1])
      if (m <= 32) { k_dilate32<<<dim3(cdiv(nw, 128), rh), 128, 0, ctx->stream>>>(T + (size_t)ry0 * nw, R + (size_t)ry0 * nw, rh, nw, m, m); LAUNCHED(); }
      else { k_dilate<<<dim3(cdiv(nw, 128), rh), 128, 0, ctx->stream>>>(T + (size_t)ry0 * nw, R + (size_t)ry0 * nw, rh, nw, m, m); LAUNCHED(); }
    }
    uint32_t* touched = nullptr; int* d_bb = nullptr;
    uint32_t* const outer_touched = ctx->touched;
    if (F.kind == COH_FILTER_SMEAR) {   // the shape of the reading scene's sprite is needed, not only its pixels
      CK(TMPGET(tmp, &touched, 4 * (size_t)rh * nw)); CK(TMPGET(tmp, &d_bb, 4 * sizeof(int)));
      CK(cudaMemsetAsync(touched, 0, 4 * (size_t)rh * nw, ctx->stream));
      ctx->touched = touched - (size_t)ry0 * nw;   // (indexed by frame row; only rows ry0 .. ry1 are visited)
    } else ctx->touched = nullptr;
    struct Restore { coh_ctx* c; uint32_t* t; ~Restore() { c->touched = t; } } restore{ctx, outer_touched};
    if (F.kind == COH_FILTER_SCENE) {
      PassArgs A{F.read0, F.read1, rbox.x0, rbox.y0, rbox.x1 - rbox.x0 + 1, rbox.y1 - rbox.y0 + 1, R, nullptr, X, true, false};
      if (render_pass(ctx, s, A)) return 1;
    } else if (render_suffix(ctx, s, F.kind == COH_FILTER_MINUS ? F.head_l1 : F.pos, fi + 1, R, X, true, rbox, true)) return 1;   // MINUS: tl scene
    Y = X;
    ctx->touched = nullptr;
    if (F.kind == COH_FILTER_SMEAR) {
      // Brush.smear (brush.ml:286-331) on a canvas around the stroke: box of bloat r r (stroke shape), one pixel of border
      const int init[4] = {INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN};
      CK(cudaMemcpyAsync(d_bb, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));   // (pageable: staged before the call returns)
      k_bits_bbox<<<dim3(cdiv(nw, 128), rh), 128, 0, ctx->stream>>>(touched, rh, nw, ry0, d_bb); LAUNCHED();
      const int cx0 = F.bx0 - F.brush_r - 2, cy0 = F.by0 - F.brush_r - 2, cw = F.bx1 - F.bx0 + 1 + 2 * F.brush_r + 4, ch = F.by1 - F.by0 + 1 + 2 * F.brush_r + 4;
      uint32_t* Cv = nullptr;
      CK(TMPGET(tmp, &Cv, 4 * (size_t)cw * ch));
      CK(TMPGET(tmp, &Y, 4 * (size_t)W * H));
      k_smear<<<1, SMEAR_THREADS, 0, ctx->stream>>>(X, Y, W, H, Cv, cx0, cy0, cw, ch, d_bb, F.bx0, F.by0, F.bx1, F.by1,
                                                  s->points + F.first2, F.count2, F.dx, F.dy, s->stamps + F.stamp_off, F.brush_r, rbox.x0, rbox.y0, rbox.x1, rbox.y1); LAUNCHED();
    }
    if (F.kind == COH_FILTER_MONOCHROME) {
      if (z_is_x) blend_flags |= 2;   // Y = monochrome of Z, taken on the fly
      else { k_monochrome<<<(unsigned)((pn + 255) / 256), 256, 0, ctx->stream>>>(X + po, X + po, pn); LAUNCHED(); }
    } else if (F.kind == COH_FILTER_BLUR) {
      // Convolve.convolve_sprite_in_shape (convolve.ml:265-296) on the canvas rows [ry0, ry1]: pixels the
      // reading scene did not render are clear, exactly like the reference's canvas outside the sprite.
      // Only the columns of shptorender are needed of the result.
      uint32_t* t1 = nullptr;
      CK(TMPGET(tmp, &t1, 4 * pn));
      if (z_is_x) { CK(TMPGET(tmp, &Y, 4 * (size_t)W * H)); }
      const int* taps = F.kernel_kind == COH_CONV_GAUSSIAN ? s->filter_taps + F.taps_off : nullptr;
      dim3 gp(cdiv(x1 - x0 + 1, 128), rh);
      k_conv_pass<<<gp, 128, 0, ctx->stream>>>(X + po, t1, W, rh, F.r, F.kernel_kind, taps, F.taps_total, 0, x0, x1); LAUNCHED();
      k_conv_pass<<<gp, 128, 0, ctx->stream>>>(t1, Y + po, W, rh, F.r, F.kernel_kind, taps, F.taps_total, 1, x0, x1); LAUNCHED();
    }
  }
  uint8_t* alpha = nullptr;
  CK(TMPGET(tmp, &alpha, (size_t)W * h));
  if (m > 0) CK(cudaMemsetAsync(R + (size_t)ry0 * nw, 0, 4 * (size_t)rh * nw, ctx->stream));   // what the reading scene left of its (bloated) update
  // R := pixels_for_normal_scene = shptorender' --- pixels_finished (render.ml:1100-1105)
  k_filter_matte<<<dim3(cdiv(nw, 4), h), 128, 0, ctx->stream>>>(T + r0, F.op + g0 * 32, F.colour, W, h, nw, alpha, R + r0); LAUNCHED();
  if (z_is_x || F.kind == COH_FILTER_SMEAR) Z = X;   // (a smear's matte is opaque all over: nothing shows through)
  else {
    CK(TMPGET(tmp, &Z, 4 * (size_t)W * H));
    CK(cudaMemset2DAsync(Z + (size_t)y0 * W + x0, 4 * (size_t)W, 0, 4 * (size_t)(x1 - x0 + 1), h, ctx->stream));
    if (render_suffix(ctx, s, F.pos, fi + 1, R, Z, true, tbox, true)) return 1;
  }
  // blend' and the composite into the accumulator; u --- ef (render.ml:1308)
  k_filter_blend<<<dim3(cdiv(W, 128), h), 128, 0, ctx->stream>>>(T + r0, alpha, Z + (size_t)y0 * W, Y ? Y + (size_t)y0 * W : nullptr, target + (size_t)y0 * W, W, h, nw, blend_flags, SG, U + r0); LAUNCHED();
  return 0;
}
// Render the scene list from leaf l0 / filter f0 to its end inside U (updated to the `u` left over)
static int render_suffix(coh_ctx* ctx, DevScene* s, int l0, int f0, uint32_t* U, uint32_t* target, bool fresh, PixBox box, bool target_zeroed) {
  const Frame& fr = ctx->fr;
  if (box.x1 < box.x0 || box.y1 < box.y0) return 0;
  // `fresh` holds until the first leaves are composited: filters alone leave the accumulator clear on what is left of U
  auto segment = [&](int a, int b) -> int {
    if (b <= a) return 0;
    PassArgs A{a, b, box.x0, box.y0, box.x1 - box.x0 + 1, box.y1 - box.y0 + 1, U, U, target, fresh, !fresh};
    if (render_pass(ctx, s, A)) return 1;
    fresh = false;
    return 0;
  };
  for (int f = f0; f < (int)s->filters.size(); f++) {
    if (segment(l0, s->filters[f].pos)) return 1;
    if (apply_filter(ctx, s, f, U, target, box, fresh)) return 1;
    l0 = s->filters[f].pos;
  }
  if (segment(l0, s->n_scene_leaves)) return 1;
  if (fresh && !target_zeroed) {   // no leaves at all: what is left of U shows nothing
    const int hh = box.y1 - box.y0 + 1;
    k_clear_in_bits<<<dim3(cdiv(fr.W, 128), hh), 128, 0, ctx->stream>>>(target + (size_t)box.y0 * fr.W, U + (size_t)box.y0 * fr.tiles_x, fr.W, hh, fr.tiles_x); LAUNCHED();
  }
  return 0;
}
static int render_filtered(coh_ctx* ctx, DevScene* s, const uint32_t* u_init, int ux, int uy, int uw, int uh) {
  const Frame& fr = ctx->fr;
  if (uw <= 0 || uh <= 0) return 0;
  // a band renders its own rows of the update; what its filters read above and below them is rendered again into the
  // filters' canvases on this context (render_pass)
  const PixBox box{std::max(ux, 0), std::max(std::max(uy, 0), fr.band_y0), std::min(ux + uw - 1, fr.W - 1), std::min(std::min(uy + uh - 1, fr.H - 1), fr.band_y1 - 1)};
  if (box.x1 < box.x0 || box.y1 < box.y0) return 0;
  const int nw = fr.tiles_x;
  const size_t nwords = (size_t)nw * fr.H;
  uint32_t* U = ctx->u_out;
  if (u_init) {
    CK(cudaMemcpyAsync(U, u_init, 4 * nwords, cudaMemcpyDeviceToDevice, ctx->stream));
    if (box.y0 > 0) CK(cudaMemsetAsync(U, 0, 4 * (size_t)box.y0 * nw, ctx->stream));
    if (box.y1 + 1 < fr.H) CK(cudaMemsetAsync(U + (size_t)(box.y1 + 1) * nw, 0, 4 * (size_t)(fr.H - 1 - box.y1) * nw, ctx->stream));
  } else { k_fill_box_bits<<<dim3(cdiv(nw, 128), fr.H), 128, 0, ctx->stream>>>(U, fr.H, nw, 0, 0, box.x0, box.y0, box.x1, box.y1); LAUNCHED(); }
  // peer framebuffers receive the finished rows in one strip copy at the end (the filter kernels do not mirror their stores)
  struct Peers { coh_ctx* c; int n; ~Peers() { c->n_peers = n; } } peers_guard{ctx, ctx->n_peers};
  ctx->n_peers = 0;
  if (render_suffix(ctx, s, 0, 0, U, ctx->fb, true, box)) return 1;
  if (s->n_leaves > s->n_front_leaves) {
    // the background list shows wherever the scene pass is not opaque (render.ml:1363-1365)
    BgPrims B; memset(&B, 0, sizeof B);
    bool prims = s->n_leaves - s->n_front_leaves <= 8;
    for (int li = s->n_front_leaves; li < s->n_leaves && prims; li++) {
      const ObjRec& o = s->h_objs[s->full.h_leaves[li]];
      prims = o.kind == K_PRIM && o.depth == 1 && o.fill.kind == 0;
      if (!prims) break;
      B.x0[B.n] = o.prim[0] + o.dx; B.y0[B.n] = o.prim[1] + o.dy; B.x1[B.n] = o.prim[2] + o.dx; B.y1[B.n] = o.prim[3] + o.dy;
      B.col[B.n] = o.fill.c0; B.pretrans[B.n] = o.pretrans; B.n++;
    }
    if (prims) {
      // plain primitives (the page, the window background): one pass over the update
      PeerFbs peers; memset(&peers, 0, sizeof peers);
      k_bg_over<<<dim3(cdiv(cdiv(fr.W, 4), 128), box.y1 - box.y0 + 1), 128, 0, ctx->stream>>>(ctx->fb, u_init, fr.W, nw, box.x0, box.y0, box.x1, box.y1, B, 0, peers); LAUNCHED();
    } else {
      uint32_t* U0 = nullptr;
      CK(DMALLOC(&U0, 4 * nwords));
      if (u_init) CK(cudaMemcpyAsync(U0, u_init, 4 * nwords, cudaMemcpyDeviceToDevice, ctx->stream));
      else { k_fill_box_bits<<<dim3(cdiv(nw, 128), fr.H), 128, 0, ctx->stream>>>(U0, fr.H, nw, 0, 0, box.x0, box.y0, box.x1, box.y1); LAUNCHED(); }
      k_not_opaque_bits<<<dim3(cdiv(nw, 4), fr.H), 128, 0, ctx->stream>>>(ctx->fb, U0, U0, fr.W, fr.H, nw); LAUNCHED();
      PassArgs A{s->n_front_leaves, s->n_leaves, box.x0, box.y0, box.x1 - box.x0 + 1, box.y1 - box.y0 + 1, U0, nullptr, ctx->fb, false, true};
      const int rc = render_pass(ctx, s, A);
      DFREE(U0);
      if (rc) return 1;
    }
  }
  for (int k = 0; k < peers_guard.n; k++)
    CK(cudaMemcpy2DAsync(ctx->peer_fb[k] + (size_t)box.y0 * fr.W + box.x0, 4 * (size_t)fr.W, ctx->fb + (size_t)box.y0 * fr.W + box.x0, 4 * (size_t)fr.W,
                         4 * (size_t)(box.x1 - box.x0 + 1), box.y1 - box.y0 + 1, cudaMemcpyDeviceToDevice, ctx->stream));
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
  if (render_frame_passes(ctx, s, A)) return 1;
  ctx->have_u = record_u;
  return 0;
}
