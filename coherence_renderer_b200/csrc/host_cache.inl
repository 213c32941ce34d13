// host_cache.inl — part of coherence_b200.cu (one translation unit; included in order): cache, object shapes, dirty regions, aliases, drag step, update-shape frames.

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
// The partial-sprite side of the cache (cache.ml:328-367) for one scene: frames served entirely from cached sprites,
// frames that had to render into them, bytes of sprite canvases and planes resident in HBM, cached objects.
int coh_cache_sprite_stats(coh_ctx* ctx, coh_scene_t scene, int64_t out[4]) {
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_cache_sprite_stats: null scene");
  size_t bytes = 0; int64_t n = 0;
  for (const SpriteEntry& e : s->sprites) if (!e.dead) { bytes += e.bytes; n++; }
  out[0] = s->sprite_hits; out[1] = s->sprite_fills; out[2] = (int64_t)bytes; out[3] = n;
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
        if (s->real_depth[k] != s->real_depth[r] + 1) continue;  // direct children only (nested groups recurse)
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
// shapeonly_of_basicshape of a filter object = the shape of its geometry (render.ml:472-474), moved with its alias
static int filter_shapes(coh_ctx* ctx, DevScene* s, const DevScene::FilterRec& F, coh_shape_t* shape, coh_shape_t* minshape, const char* who) {
  coh_shape_t fs = 0, fm = 0;
  if (F.geom_sub) {   // the geometry is an object of its own, kept as a scene in canvas coordinates; the minshape is not kept
    *shape = 0; *minshape = 0;
    if (subscene_shape(ctx, F.geom_sub, &fs)) return 1;
    int rc = coh_shape_translate(ctx, fs, F.gcx0 + F.dx, F.gcy0 + F.dy, shape);
    coh_shape_free(ctx, fs);
    return rc;
  }
  if (F.kind == COH_FILTER_SMEAR) {   // the stroke's dummy brush: boxes around its stamp points
    *shape = 0; *minshape = 0;
    const int x0 = F.bx0 - F.dx, y0 = F.by0 - F.dy, x1 = F.bx1 - F.dx, y1 = F.by1 - F.dy;
    const int wx0 = floordiv(x0, 32) * 32, nw = (x1 - wx0) / 32 + 1, n_rows = y1 - y0 + 1;
    uint32_t* bits = nullptr;
    CK(DMALLOC(&bits, 4 * (size_t)nw * n_rows));
    CK(cudaMemsetAsync(bits, 0, 4 * (size_t)nw * n_rows, ctx->stream));
    const int side = 2 * F.brush_r + 1;
    k_stamp_boxes_to_bits<<<cdiv(F.count * side, 256), 256, 0, ctx->stream>>>(s->points + F.first, F.count, F.brush_r, y0, n_rows, wx0, nw, bits); LAUNCHED();
    int rc = shape_from_bits(ctx, bits, y0, n_rows, wx0, nw, &fs);
    DFREE(bits);
    if (!rc) rc = coh_shape_translate(ctx, fs, F.dx, F.dy, shape);
    coh_shape_free(ctx, fs);
    return rc;
  }
  if (shapes_from_device_edges(ctx, s->edges + F.first, F.count, F.winding, F.bx0 - F.dx, F.by0 - F.dy, F.bx1 - F.dx, F.by1 - F.dy, &fs, &fm, who)) return 1;
  if (!F.dx && !F.dy) { *shape = fs; *minshape = fm; return 0; }
  int rc = coh_shape_translate(ctx, fs, F.dx, F.dy, shape) || coh_shape_translate(ctx, fm, F.dx, F.dy, minshape);
  coh_shape_free(ctx, fs); coh_shape_free(ctx, fm);
  return rc;
}
static DevScene::FilterRec* filter_of_abi(DevScene* s, int obj_index) {
  for (DevScene::FilterRec& F : s->filters) if (F.abi == obj_index) return &F;
  return nullptr;
}
int coh_scene_object_shape(coh_ctx* ctx, coh_scene_t scene, int32_t obj_index, coh_shape_t* shape, coh_shape_t* minshape) {
  CK(cudaSetDevice(ctx->device));
  DevScene* s = (DevScene*)scene;
  if (!s) FAIL("coh_scene_object_shape: null scene");
  if (const DevScene::FilterRec* F = filter_of_abi(s, obj_index)) return filter_shapes(ctx, s, *F, shape, minshape, "coh_scene_object_shape");
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
  if (DevScene::FilterRec* F = filter_of_abi(s, obj_index)) {
    // a lens moves: its kept planes are re-made at the next frame (the key holds the offset)
    F->dx += dx; F->dy += dy; F->bx0 += dx; F->bx1 += dx; F->by0 += dy; F->by1 += dy;
    return 0;
  }
  if (obj_index < 0 || obj_index >= (int)s->rec_of_abi.size() || s->rec_of_abi[obj_index] < 0) FAIL("coh_scene_translate_object: no such object");
  const int r = s->rec_of_abi[obj_index];
  const int last = s->h_objs[r].kind == K_GROUP ? s->group_last[r] : r;
  // the kept shapes of every group around the moved object (or group) are stale; a moved group's own entry stays
  // valid (it is corrected through group_off on the way out)
  for (auto git = s->group_shape.begin(); git != s->group_shape.end();) {
    const int g = git->first;
    if (g < r && s->group_last[g] >= last) { free_shape(ctx, git->second.shape); git = s->group_shape.erase(git); }
    else ++git;
  }
  for (int k = r; k <= last; k++) {
    ObjRec& o = s->h_objs[k];
    if (o.kind == K_GROUP) { s->group_off[k].x += dx; s->group_off[k].y += dy; continue; }
    o.dx += dx; o.dy += dy; o.bx0 += dx; o.bx1 += dx; o.by0 += dy; o.by1 += dy;
  }
  s->full.items_for_W = -1; s->sp.items_for_W = -1;    // the item-pool bounds depend on the boxes
  s->full.bins_valid = false; s->sp.bins_valid = false;  // ... and so do the cell lists
  if (s->n_leaves > 0) { k_move_leaves<<<cdiv(s->n_leaves, 256), 256, 0, ctx->stream>>>(s->objs, s->leaf_box, s->leaves, s->n_leaves, r, last, dx, dy, 1); LAUNCHED(); }
  if (s->sp.n > 0) {
    // the list with cached objects collapsed: boxes of the leaves it shares follow the records; a cached object moved as
    // a whole takes its sprite leaf along (the alias reads the cached sprite translated, cache.ml:400-405); one whose
    // member moved on its own is no longer the object that was cached
    // (the first sprite leaf that moves along rides in the same launch: its record lies outside [r, last])
    int along = -1;
    for (const SpriteEntry& e : s->sprites) if (e.grp >= r && e.grp <= last && (e.leaf_rec < r || e.leaf_rec > last)) { along = e.leaf_rec; break; }
    k_move_leaves<<<cdiv(s->sp.n, 256), 256, 0, ctx->stream>>>(s->objs, s->sp.leaf_box, s->sp.leaves, s->sp.n, r, last, dx, dy, 0, along); LAUNCHED();
    for (SpriteEntry& e : s->sprites) {
      if (e.grp >= r && e.grp <= last) {
        ObjRec& o = s->h_objs[e.leaf_rec];
        o.dx += dx; o.dy += dy; o.bx0 += dx; o.bx1 += dx; o.by0 += dy; o.by1 += dy;
        if (e.leaf_rec != along) { k_move_leaves<<<cdiv(s->sp.n, 256), 256, 0, ctx->stream>>>(s->objs, s->sp.leaf_box, s->sp.leaves, s->sp.n, e.leaf_rec, e.leaf_rec, dx, dy, 1); LAUNCHED(); }
      } else if (r > e.grp && r <= s->group_last[e.grp]) e.dead = true;
    }
  }
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
  int rc = render_frame_passes(ctx, s, A);
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
    if (filter_shapes(ctx, s, F, &fs, &fm, "coh_dirty_filter")) return 1;
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
  const DevScene::FilterRec* Fm = filter_of_abi(s, obj_index);
  if (!Fm && (obj_index < 0 || obj_index >= (int)s->rec_of_abi.size() || s->rec_of_abi[obj_index] < 0)) FAIL("coh_scene_drag_object: no such object");
  // Fill.Plain objects: plaindirty; groups, fancy fills, brush strokes, Convolved objects, filters: alldirty (render.ml:1396-1400)
  bool plain = false, prim = false;
  coh_shape_t so = 0, mo = 0;
  const DevShape* kept = nullptr; int kox = 0, koy = 0;
  if (Fm) { if (filter_shapes(ctx, s, *Fm, &so, &mo, "coh_scene_drag_object")) return 1; }
  else {
    const int r = s->rec_of_abi[obj_index];
    const ObjRec& o = s->h_objs[r];
    plain = (o.kind == K_PATH || o.kind == K_CPG) && o.fill.kind == 0;
    prim = o.kind == K_PRIM;
    // a group whose shape is kept with the scene: the span set is read in place through the group's offset (what
    // object_shape_rec would hand out as a translated copy: two allocations, a copy and a launch per step)
    if (o.kind == K_GROUP && ctx->usecache) {
      auto git = s->group_shape.find(r);
      if (git != s->group_shape.end() && git->second.shape) {
        kept = git->second.shape; kox = s->group_off[r].x - git->second.offx; koy = s->group_off[r].y - git->second.offy;
        ctx->shphit++;
      }
    }
    if (!kept && object_shape_rec(ctx, s, r, &so, &mo)) return 1;   // served by the cache after the first step
  }
  if (coh_scene_translate_object(ctx, scene, obj_index, dx, dy)) return 1;
  const Frame& fr = ctx->fr;
  const int nw = fr.tiles_x;
  const size_t nwords = (size_t)nw * fr.H;
  if (!ctx->u_init) CK(DMALLOC(&ctx->u_init, sizeof(uint32_t) * nwords));
  uint32_t* U = ctx->u_init;
  CK(cudaMemsetAsync(U, 0, 4 * nwords, ctx->stream));
  const DevShape* S = kept ? kept : (const DevShape*)so; DevShape* M = (DevShape*)mo;
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
      // shp_o ||| shp_n: the old position (offset 0) and the new one (the same span set read through (dx, dy)) in one launch
      k_spans_to_bits2<<<cdiv(fr.H * 32, 256), 256, 0, ctx->stream>>>(S->row_ptr, S->spans, S->y0 + koy, -kox, S->y0 + koy + dy, -(kox + dx), S->n_rows, fr.H, nw, U); LAUNCHED();
    }
    bb[0] = std::max(0, S->bx0 + kox + std::min(dx, 0)); bb[1] = std::max(0, S->by0 + koy + std::min(dy, 0));
    bb[2] = std::min(fr.W - 1, S->bx1 + kox + std::max(dx, 0)); bb[3] = std::min(fr.H - 1, S->by1 + koy + std::max(dy, 0));
  }
  coh_shape_free(ctx, so); coh_shape_free(ctx, mo);
  if (dirty_bbox) for (int k = 0; k < 4; k++) dirty_bbox[k] = bb[k];
  ctx->have_u = false;
  if (bb[2] < bb[0] || bb[3] < bb[1]) return 0;
  const bool record_u = (flags & COH_RENDER_RECORD_U) != 0;
  int rc;
  if (!s->filters.empty()) { rc = render_filtered(ctx, s, U, bb[0], bb[1], bb[2] - bb[0] + 1, bb[3] - bb[1] + 1); ctx->have_u = !rc; return rc; }
  PassArgs A{0, s->n_leaves, bb[0], bb[1], bb[2] - bb[0] + 1, bb[3] - bb[1] + 1, U, record_u ? ctx->u_out : nullptr, ctx->fb, true, false};
  rc = render_frame_passes(ctx, s, A);
  ctx->have_u = record_u && !rc;
  return rc;
}

// Convolved (kernel, Group members): shape = bloat r r (union of the members' shapes), minshape null (render.ml:536-555);
// sprite = convolve_sprite kernel (the group's sprite) (render.ml:1023-1052).  The members' scene is rendered once
// over the whole twice-bloated box into a canvas whose origin is moved to the frame origin by aliasing the members.
// Union of the shapes of a sub-scene's direct members (filters count with their geometry's shape, render.ml:472-474), with the
// cache switched off as the reference does while it computes a group's shape (render.ml:476-496)
static int subscene_shape(coh_ctx* ctx, DevScene* ss, coh_shape_t* out) {
  const bool saved = ctx->usecache; ctx->usecache = false;
  coh_shape_t gs = 0; int rc = 0;
  for (int k = 1; k < (int)ss->real_depth.size() && !rc; k++) {
    if (ss->real_depth[k] != 1) continue;   // direct members
    coh_shape_t ms = 0, mm = 0, un = 0;
    rc = object_shape_rec(ctx, ss, k, &ms, &mm) || coh_shape_union(ctx, gs, ms, &un);
    coh_shape_free(ctx, gs); coh_shape_free(ctx, ms); coh_shape_free(ctx, mm);
    gs = un;
  }
  for (const DevScene::FilterRec& F : ss->filters) {
    if (rc) break;
    coh_shape_t ms = 0, mm = 0, un = 0;
    rc = filter_shapes(ctx, ss, F, &ms, &mm, "coh_scene_create (group)") || coh_shape_union(ctx, gs, ms, &un);
    coh_shape_free(ctx, gs); coh_shape_free(ctx, ms); coh_shape_free(ctx, mm);
    gs = un;
  }
  ctx->usecache = saved;
  *out = gs;
  return rc;
}
// Alias the direct members of a sub-scene (objs: the records it was created from) by (dx, dy)
static int subscene_translate(coh_ctx* ctx, DevScene* ss, const coh_object* objs, int n, int dx, int dy) {
  int depth = 0;
  for (int t = 0; t < n; t++) {
    const int kind = objs[t].kind;
    if (kind == COH_OBJ_GROUP_END) { depth--; continue; }
    if (depth == 0 && ((t < (int)ss->rec_of_abi.size() && ss->rec_of_abi[t] >= 0) || filter_of_abi(ss, t)))
      if (coh_scene_translate_object(ctx, (coh_scene_t)ss, t, dx, dy)) return 1;
    if (kind == COH_OBJ_GROUP_BEGIN && !objs[t].convolve) depth++;
  }
  return 0;
}
// Render a sub-scene, whose frame is a canvas of nw words x h rows, into `A` (cleared by the caller)
static int subscene_render(coh_ctx* ctx, DevScene* ss, int nw, int h, uint32_t* A) {
  const int w = nw * 32;
  const Frame saved = ctx->fr;
  ctx->fr.W = w; ctx->fr.H = h; ctx->fr.band_y0 = 0; ctx->fr.band_y1 = h; ctx->fr.tiles_x = nw; ctx->fr.cells_y = cdiv(h, CELL_H); ctx->fr.ctx0 = 0; ctx->fr.cntx = nw;
  int rc;
  if (ss->filters.empty()) {
    PassArgs pa{0, ss->n_leaves, 0, 0, w, h, nullptr, nullptr, A, true, false};
    rc = render_pass(ctx, ss, pa);
  } else {
    // filter passes work on the context's frame: the canvas stands in for it
    uint32_t* saved_fb = ctx->fb; uint32_t* saved_u = ctx->u_out; const int saved_peers = ctx->n_peers;
    uint32_t* Uc = nullptr;
    if (DMALLOC(&Uc, 4 * (size_t)nw * h) != cudaSuccess) { ctx->fr = saved; FAIL("sub-scene render: out of memory"); }
    ctx->fb = A; ctx->u_out = Uc; ctx->n_peers = 0;
    rc = render_filtered(ctx, ss, nullptr, 0, 0, w, h);
    ctx->fb = saved_fb; ctx->u_out = saved_u; ctx->n_peers = saved_peers;
    DFREE(Uc);
  }
  ctx->fr = saved;
  return rc;
}
static int realize_convolved_group(coh_ctx* ctx, DevScene* s, const ObjRec& o, ConvGroup& cg) {
  DevScene* ss = cg.sub;
  const int nw = o.cv_nw, h = o.cv_h, w = nw * 32;
  const size_t nwords = (size_t)nw * h, npx = (size_t)w * h;
  uint32_t *S = nullptr, *A = nullptr, *X = nullptr; int* d_taps = nullptr;
  CK(DMALLOC(&S, 4 * nwords)); CK(DMALLOC(&A, 4 * npx)); CK(DMALLOC(&X, 4 * npx));
  CK(cudaMemsetAsync(S, 0, 4 * nwords, ctx->stream)); CK(cudaMemsetAsync(A, 0, 4 * npx, ctx->stream));
  // (1) the group's shape
  {
    coh_shape_t gs = 0;
    if (subscene_shape(ctx, ss, &gs)) return 1;
    if (gs) {
      const DevShape* G = (const DevShape*)gs;
      k_spans_to_bits<<<cdiv(h, 128), 128, 0, ctx->stream>>>(G->row_ptr, G->spans, G->y0, G->n_rows, o.cv_y0, h, o.cv_x0, nw, S); LAUNCHED();
      coh_shape_free(ctx, gs);
    }
  }
  uint32_t* convS = s->conv_bits + o.cv_bits;
  if (cg.r > 0) { k_dilate<<<dim3(cdiv(nw, 128), h), 128, 0, ctx->stream>>>(S, convS, h, nw, cg.r, cg.r); LAUNCHED(); }
  else CK(cudaMemcpyAsync(convS, S, 4 * nwords, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemsetAsync(convS + nwords, 0, 4 * nwords, ctx->stream));
  // (2) the group's sprite: the members' scene rendered with the canvas as its frame
  if (subscene_translate(ctx, ss, cg.members, cg.n_members, -o.cv_x0, -o.cv_y0)) return 1;
  if (subscene_render(ctx, ss, nw, h, A)) return 1;
  if (cg.r == 0) {   // no kernel: the canvas is the sprite
    CK(cudaMemcpyAsync(s->conv_px + o.cv_px, A, 4 * npx, cudaMemcpyDeviceToDevice, ctx->stream));
    if (check_error_flag(ctx, "coh_scene_create (group holding filters)")) return 1;
    DFREE(S); DFREE(A); DFREE(X);
    coh_scene_free(ctx, (coh_scene_t)ss); cg.sub = nullptr;
    return 0;
  }
  // (3) Convolve.convolve_sprite on the canvas: X pass, Y pass
  std::vector<int> taps; int total = 0;
  if (cg.kind == COH_CONV_GAUSSIAN) {  // Convolve.mkgaussian r (convolve.ml:60-70)
    for (int i = -cg.r; i <= cg.r; i++) {
      double xr = (double)i / (double)cg.r, yr = 0. / (double)cg.r;
      double gg = exp(-(xr * xr + yr * yr)) / 2.;
      int v = (int)((double)(4 * cg.r * cg.r) * gg + 0.5);
      taps.push_back(v); total += v;
    }
    CK(DMALLOC(&d_taps, sizeof(int) * taps.size()));
    CK(cudaMemcpyAsync(d_taps, taps.data(), sizeof(int) * taps.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  dim3 gp(cdiv(w, 128), h);
  k_conv_pass<<<gp, 128, 0, ctx->stream>>>(A, X, w, h, cg.r, cg.kind, d_taps, total, 0); LAUNCHED();
  k_conv_pass<<<gp, 128, 0, ctx->stream>>>(X, s->conv_px + o.cv_px, w, h, cg.r, cg.kind, d_taps, total, 1); LAUNCHED();
  if (check_error_flag(ctx, "coh_scene_create (Convolved group)")) return 1;
  DFREE(S); DFREE(A); DFREE(X); DFREE(d_taps);
  coh_scene_free(ctx, (coh_scene_t)ss); cg.sub = nullptr;
  return 0;
}
