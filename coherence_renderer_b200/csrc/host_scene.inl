// host_scene.inl — part of coherence_b200.cu (one translation unit; included in order): scene creation (flattened renderobject list -> device records, K1 edge binning, Convolved pre-pass) and the framebuffer.

// ---------------------------------------------------------------------------------------
// Scenes and rendering
// ---------------------------------------------------------------------------------------
int coh_scene_free(coh_ctx* ctx, coh_scene_t h) {
  CK(cudaSetDevice(ctx->device));
  DevScene* s = (DevScene*)h;
  if (!s) return 0;
  DFREE(s->objs); DFREE(s->leaves); DFREE(s->leaf_box); DFREE(s->edges); DFREE(s->points); DFREE(s->stamps);
  DFREE(s->rowedge_ptr); DFREE(s->rowedge_idx); DFREE(s->brush_ranges); DFREE(s->conv_bits); DFREE(s->conv_px); DFREE(s->attr); DFREE(s->filter_taps);
  for (DevScene::FilterRec& f : s->filters) { DFREE(f.SG); DFREE(f.CG); DFREE(f.op); if (f.geom_sub) coh_scene_free(ctx, (coh_scene_t)f.geom_sub); }
  for (auto& g : s->group_shape) free_shape(ctx, g.second.shape);
  free_binset(ctx, s->full.bins); free_binset(ctx, s->sp.bins);
  DFREE(s->sp.leaves); DFREE(s->sp.leaf_box);
  for (SpriteEntry& e : s->sprites) {
    if (e.ev_pending) cudaEventSynchronize(e.ev);
    DFREE(e.valid); DFREE(e.d_missing);
    if (e.h_missing) cudaFreeHost(e.h_missing);
    if (e.ev) cudaEventDestroy(e.ev);
  }
  delete s;
  return 0;
}

// brush.ml:60-92: alpha of the Gaussian stamp of white at `opacity`.
// The Gaussian of a stamp (brush.ml:60-92) is a function of the radius alone; the opacity only dissolves it.  Scenes use
// few radii and many opacities (C3: 18 radii, ~2300 stamps), so the exponentials are kept per radius for the duration
// of a call (`gauss`, keyed by the bits of the radius): same expression, same roundings, computed once.
typedef std::map<uint64_t, std::vector<uint8_t>> StampGauss;
static void brush_stamp(double radius, double opacity, std::vector<uint8_t>& out, int& r_out, StampGauss* gauss = nullptr) {
  int intopacity = (int)(opacity * 255.), intr = (int)ceil(radius);
  int size = 2 * intr + 1;
  r_out = intr;
  size_t base = out.size();
  out.resize(base + (size_t)size * size);
  uint32_t white = 0xFFFFFFFFu;
  uint32_t c1 = px_dissolve(white, intopacity);
  StampGauss local;
  uint64_t rbits; memcpy(&rbits, &radius, sizeof rbits);
  std::vector<uint8_t>& g = (gauss ? *gauss : local)[rbits];
  if (g.empty()) {
    g.resize((size_t)size * size);
    for (int y = 0; y < size; y++)
      for (int x = 0; x < size; x++) {
        double xp = (double)(x - intr), yp = (double)(y - intr), rr = radius / 2.;
        double v = 255. * exp(-((xp / rr) * (xp / rr) + (yp / rr) * (yp / rr)));
        g[(size_t)y * size + x] = (uint8_t)(int)(v * 1.);   // 0 .. 255
      }
  }
  uint8_t of_v[256];   // px_dissolve (c1, vi) >> 24 for every vi
  for (int vi = 0; vi < 256; vi++) of_v[vi] = (uint8_t)(px_dissolve(c1, vi) >> 24);
  for (size_t k = 0; k < (size_t)size * size; k++) out[base + k] = of_v[g[k]];
}

// Convolved (kernel, Group members): the members are a scene of their own (realised in host_cache.inl, after the
// render passes and the shape functions it needs)
struct ConvGroup { int rec, kind, r; DevScene* sub; const coh_object* members; int n_members; };
static int realize_convolved_group(coh_ctx* ctx, DevScene* s, const ObjRec& o, ConvGroup& cg);
static int subscene_translate(coh_ctx* ctx, DevScene* ss, const coh_object* objs, int n, int dx, int dy);
// Ownership of edges / brush points by objects, as ranges (see coh_scene_create).
struct OwnerRanges {
  std::vector<int4> ranges;   // first, count, record, -
  bool monotone = true; long long last_end = 0;
  void add(int first, int count, int rec) {
    if (count <= 0) return;
    if (first < last_end) monotone = false;
    last_end = std::max(last_end, (long long)first + count);
    ranges.push_back(make_int4(first, count, rec, 0));
  }
  bool overlapping() const {
    if (monotone) return false;   // every range began at or after the end of all earlier ones
    std::vector<int4> r(ranges);
    std::sort(r.begin(), r.end(), [](const int4& a, const int4& b) { return a.x < b.x; });
    for (size_t k = 1; k < r.size(); k++) if ((long long)r[k - 1].x + r[k - 1].y > r[k].x) return true;
    return false;
  }
};
// per-element owner array on the device (-1: no owner) from the ranges
static int expand_owners(coh_ctx* ctx, const OwnerRanges& own, int n_elems, int** d_owner) {
  CK(DMALLOC(d_owner, sizeof(int) * (size_t)std::max(n_elems, 1)));
  CK(cudaMemsetAsync(*d_owner, 0xFF, sizeof(int) * (size_t)std::max(n_elems, 1), ctx->stream));
  if (own.ranges.empty()) return 0;
  int4* d_rg = nullptr;
  CK(DMALLOC(&d_rg, sizeof(int4) * own.ranges.size()));
  CK(cudaMemcpyAsync(d_rg, own.ranges.data(), sizeof(int4) * own.ranges.size(), cudaMemcpyHostToDevice, ctx->stream));
  k_fill_owner<<<cdiv((int)own.ranges.size() * 32, 256), 256, 0, ctx->stream>>>(d_rg, (int)own.ranges.size(), *d_owner); LAUNCHED();
  CK(cudaStreamSynchronize(ctx->stream));   // (the range list is a local of the caller's frame)
  DFREE(d_rg);
  return 0;
}
// Bounds of the elements of every object: a / b = the edge lists of a path, a filter's geometry or the operands of a
// CPG; for a brush stroke a = the box of its points.  Ranges that are out of bounds are skipped here — the object
// loop rejects them before it looks at the result.
struct ElemBounds { EdgeBox a, b; };
static void element_bounds(const coh_object* objs, int n_objs, const int32_t* edges, int n_edges, const int32_t* points, int n_points,
                           std::vector<ElemBounds>& out) {
  out.resize((size_t)std::max(n_objs, 1));
  auto in_range = [](int first, int count, int n) { return first >= 0 && count >= 0 && (int64_t)first + count <= n; };
  auto one = [&](int i) {
    const coh_object& c = objs[i];
    ElemBounds& r = out[i];
    if (c.kind == COH_OBJ_BRUSH) {
      if (!in_range(c.first, c.count, n_points)) return;
      EdgeBox b{INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
      const int32_t* p = points + 2 * (size_t)c.first;
      for (int k = 0; k < c.count; k++) {
        b.xmin = std::min(b.xmin, p[2 * k]); b.xmax = std::max(b.xmax, p[2 * k]);
        b.ymin = std::min(b.ymin, p[2 * k + 1]); b.ymax = std::max(b.ymax, p[2 * k + 1]);
      }
      r.a = b;
    } else if (c.kind == COH_OBJ_PATH || c.kind == COH_OBJ_CPG || c.kind == COH_OBJ_FILTER) {
      if (in_range(c.first, c.count, n_edges)) r.a = edge_bounds(edges + 4 * (size_t)c.first, c.count);
      if (c.kind == COH_OBJ_CPG && in_range(c.first2, c.count2, n_edges)) r.b = edge_bounds(edges + 4 * (size_t)c.first2, c.count2);
    }
  };
  const long long work = (long long)n_edges + n_points;
  unsigned nt = work > 400000 ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
  if (nt <= 1) { for (int i = 0; i < n_objs; i++) one(i); return; }
  std::atomic<int> next(0);
  const int chunk = 256;
  auto worker = [&]() { for (;;) { const int i0 = next.fetch_add(chunk); if (i0 >= n_objs) return; for (int i = i0; i < std::min(n_objs, i0 + chunk); i++) one(i); } };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
}

int coh_scene_create(coh_ctx* ctx, const coh_object* objs, int32_t n_objs, int32_t n_background, const int32_t* edges,
                     int32_t n_edges, const int32_t* points, int32_t n_points, coh_scene_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  // option "trace_create": phase times of this call on stderr (synchronises at every mark)
  auto t_last = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!(ctx->opt_ab & 2)) return;
    cudaStreamSynchronize(ctx->stream);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "coh_scene_create: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  if (n_objs < 0 || n_background < 0 || n_background > n_objs) FAIL("scene: bad object counts");
  // The scene list and the (pages @ background) list are each wrapped in an implicit root group:
  // render_frame renders them separately over the same update and composites the two results
  // with `over` (render.ml:1357-1365), which is exactly what two sibling groups do in one walk.
  // the scene is built in place: records and leaf list are the scene's own vectors (no copy of 10^5 records at the end)
  std::unique_ptr<DevScene> holder(new DevScene());
  DevScene* s = holder.get();
  std::vector<ObjRec>& recs = s->h_objs;
  std::vector<int>& leaves = s->h_leaves;
  // The brush points of a large scene (C3: 88 MB of pageable host memory, 9 ms) go up on a stream of their own from a helper
  // thread while this thread walks the objects.  (Declared after `holder`: joined before the scene can go away.)
  struct PointUpload {
    std::thread th; cudaStream_t st = nullptr; cudaEvent_t ev = nullptr; cudaError_t err = cudaSuccess;
    int2** slot = nullptr; cudaStream_t main = nullptr; bool taken = false;   // a call that fails gives the early allocation back
    ~PointUpload() {
      if (th.joinable()) th.join();
      if (!taken && slot && *slot) { cudaFreeAsync(*slot, main); *slot = nullptr; }
      if (ev) cudaEventDestroy(ev);
      if (st) cudaStreamDestroy(st);
    }
  } up;
  if (n_points >= (1 << 20)) {
    const size_t bytes = sizeof(int2) * (size_t)n_points;
    CK(DMALLOC(&s->points, bytes));
    up.slot = &s->points; up.main = ctx->stream;
    CK(cudaStreamCreateWithFlags(&up.st, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&up.ev, cudaEventDisableTiming));
    CK(cudaEventRecord(up.ev, ctx->stream));   // the allocation is ordered on the context's stream
    CK(cudaStreamWaitEvent(up.st, up.ev, 0));
    PointUpload* u = &up; const int dev = ctx->device; int2* dst = s->points;
    up.th = std::thread([u, dev, dst, points, bytes] {
      u->err = cudaSetDevice(dev);
      if (u->err == cudaSuccess) u->err = cudaMemcpyAsync(dst, points, bytes, cudaMemcpyHostToDevice, u->st);
      if (u->err == cudaSuccess) u->err = cudaStreamSynchronize(u->st);   // resident before this thread is joined
    });
  }
  std::vector<uint8_t> stamps;
  std::map<std::pair<uint64_t, int>, std::pair<int, int>> stamp_cache;   // (radius, integer opacity) -> (offset, r)
  StampGauss stamp_gauss;
  std::vector<int> open;  // indices (into recs) of open groups
  // A Group that is the first member of its list and is composited with plain Over goes under an accumulator that is
  // still clear, and `over clear s = s` exactly (colour.ml:314-316): its members can composite straight into the
  // parent's accumulator — same pixels, same u — so such groups vanish from the members' ancestor chains (the lion of
  // examples.ml:174-180 in front of a scene is one).  Filter objects do not count as members here: a filter finishes
  // the whole of its shape (render.ml:1120-1121), so every pixel still in `u` after it has a clear accumulator, and the
  // pass that follows filters alone starts fresh (render_suffix).
  std::vector<int> eff_open;      // the ancestor chain the walker sees: open groups that are not dissolved into their parent
  std::vector<int> n_children;    // per open group: members seen so far
  std::vector<char> open_flat;    // per open group: dissolved into its parent
  // Owning object of every edge / brush point: kept as (first, count, record) ranges here and expanded on the device
  // (k_fill_owner); ranges of different objects may not overlap.
  OwnerRanges edge_own, point_own;
  // Bounds of every object's elements (edge lists, brush points) are independent of everything else in the scene:
  // computed up front by a few host threads when the scene is large.
  std::vector<ElemBounds> pre_bounds;
  element_bounds(objs, n_objs, edges, n_edges, points, n_points, pre_bounds);
  mark("element bounds");
  struct ConvItem { int rec, kind, r; };
  std::vector<ConvItem> conv_list;
  std::vector<ConvGroup> conv_groups;   // Convolved (kernel, Group members): the members are a scene of their own
  struct SubScenes { coh_ctx* ctx; std::vector<ConvGroup>* v; ~SubScenes() { for (ConvGroup& g : *v) if (g.sub) coh_scene_free(ctx, (coh_scene_t)g.sub); } } sub_guard{ctx, &conv_groups};
  size_t conv_words = 0, conv_pixels = 0;
  long long total_rows = 0, total_brush_rows = 0;
  ObjRec root; memset(&root, 0, sizeof root);
  root.kind = K_GROUP; root.pretrans = -1; root.depth = 0; root.flags = OF_ROOT_SCENE;
  recs.reserve((size_t)n_objs + 4); leaves.reserve((size_t)n_objs);
  recs.push_back(root); open.push_back(0);
  eff_open.push_back(0); n_children.push_back(0); open_flat.push_back(0);
  std::vector<int> rec_of_abi((size_t)std::max(n_objs, 1), -1);
  std::vector<int> group_last;
  std::vector<int> real_depth;   // per record: number of enclosing groups in the scene as given (roots included)
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
      eff_open[0] = open[0]; n_children[0] = 0;
    }
    const coh_object& c = objs[i];
    int skip_to = -1;
    if (c.kind == COH_OBJ_GROUP_END) {
      if (open_reading.empty()) FAIL("scene: GROUP_END without GROUP_BEGIN");
      const int rd = open_reading.back(); open_reading.pop_back();
      if (rd >= 0) { reading[rd].second = (int)leaves.size(); cur_reading = -1; continue; }
      if (open.size() <= 1) FAIL("scene: GROUP_END without GROUP_BEGIN");
      group_last.resize(recs.size(), -1);
      group_last[open.back()] = (int)recs.size() - 1;
      open.pop_back();
      if (!open_flat.back()) eff_open.pop_back();
      open_flat.pop_back(); n_children.pop_back();
      continue;
    }
    if (c.kind == COH_OBJ_GROUP_BEGIN && c.filter_kind == COH_FILTER_READING_SCENE) {
      // members become direct members of the root list of their own pass (render.ml:1091 renders the list)
      if (open.size() != 1 || cur_reading >= 0 || i >= n_objs - n_background) FAIL("scene: reading-scene groups must be top-level members of the scene list");
      if (n_scene_leaves < 0) n_scene_leaves = (int)leaves.size();
      open_reading.push_back(i); cur_reading = i;
      n_children[0] = 0;   // a list of its own, rendered from a clear accumulator (render.ml:1091)
      reading[i] = std::make_pair((int)leaves.size(), (int)leaves.size());
      continue;
    }
    if (n_scene_leaves >= 0 && cur_reading < 0 && i < n_objs - n_background) FAIL("scene: reading-scene groups must come after every ordinary scene object");
    if (c.kind == COH_OBJ_FILTER) {
      if (open.size() != 1 || cur_reading >= 0 || i >= n_objs - n_background) FAIL("scene: filter objects must be top-level members of the scene list");
      if (c.filter_kind < COH_FILTER_HOLE || c.filter_kind > COH_FILTER_SMEAR) FAIL("scene: bad filter kind");
      const bool smear = c.filter_kind == COH_FILTER_SMEAR;
      bool geom_next = !smear && c.cpg_op == COH_GEOM_NEXT;
      int gj = i;   // last record of the geometry object, when it follows the filter
      const coh_object* own = &c;   // the record that describes a path geometry
      if (geom_next && i + 1 < n_objs - n_background && objs[i + 1].kind == COH_OBJ_PATH && !objs[i + 1].convolve) {
        // a plain path (or stroked-path outline) as the following object is the same thing as the filter's own path:
        // Polygon.polygon_sprite samples EVERY pixel it is given (render.ml:1018, polygon.ml:729-746), which the path
        // route reproduces and a scene render (interior pixels filled, not sampled) would not
        own = &objs[i + 1]; geom_next = false; gj = i + 1;
      }
      if (geom_next) {
        if (i + 1 >= n_objs - n_background || objs[i + 1].kind == COH_OBJ_GROUP_END || objs[i + 1].kind == COH_OBJ_FILTER ||
            (objs[i + 1].kind == COH_OBJ_GROUP_BEGIN && objs[i + 1].filter_kind == COH_FILTER_READING_SCENE))
          FAIL("scene: a filter with COH_GEOM_NEXT is followed by its geometry object");
        gj = i + 1;
        // sprite_of_cpg asked for ALL the pixels of the object combines the operands' SAMPLED alphas also inside their
        // minshapes (render.ml:867-981), where a scene render takes the fill: wrap the CPG in a group to get the latter
        if (objs[gj].kind == COH_OBJ_CPG) FAIL("scene: a CPG as a filter's geometry itself is not supported (make it the member of a group)");
        if (objs[gj].kind == COH_OBJ_GROUP_BEGIN) {
          int nest = 1;
          for (gj = i + 2; gj < n_objs - n_background; gj++) {
            if (objs[gj].kind == COH_OBJ_GROUP_BEGIN) nest++;
            else if (objs[gj].kind == COH_OBJ_GROUP_END && --nest == 0) break;
          }
          if (gj >= n_objs - n_background) FAIL("scene: unterminated group");
        }
        for (int k = i + 1; k <= gj; k++) {
          const coh_object& m = objs[k];
          if (m.kind == COH_OBJ_FILTER) FAIL("scene: a filter's geometry may not hold filters");
          if (m.kind != COH_OBJ_GROUP_BEGIN && m.kind != COH_OBJ_GROUP_END && m.kind != COH_OBJ_PRIMITIVE && m.fill_kind != COH_FILL_PLAIN)
            FAIL("scene: filter geometry with a fancy fill is not supported yet");
        }
      } else if (smear) {   // geometry = the stroke's dummy: stamp points; smear points beside them (both in the points array)
        if (c.first < 0 || c.count < 0 || (int64_t)c.first + c.count > n_points || c.first2 < 0 || c.count2 < 0 || (int64_t)c.first2 + c.count2 > n_points) FAIL("scene: point range out of bounds");
        if (!(c.brush_radius >= 0. && c.brush_radius <= 64.) || !(c.brush_opacity >= 0. && c.brush_opacity <= 1.)) FAIL("scene: brush radius/opacity out of range (a smear brush has a radius of at most 64)");
      } else {
        if (own->first < 0 || own->count < 0 || (int64_t)own->first + own->count > n_edges) FAIL("scene: edge range out of bounds");
        if (own->winding != COH_NONZERO && own->winding != COH_EVENODD) FAIL("scene: bad winding rule");
        if (own->sprite_winding < 0 || own->sprite_winding > 2) FAIL("scene: bad sprite winding rule");
        if (own->fill_kind != COH_FILL_PLAIN) FAIL("scene: filter geometry with a fancy fill is not supported yet");
      }
      DevScene::FilterRec f; memset(&f, 0, sizeof f);
      f.abi = i; f.pos = (int)leaves.size(); f.kind = c.filter_kind; f.first = own->first; f.count = own->count; f.winding = own->winding; f.colour = own->colour0;
      f.aa_winding = own->sprite_winding ? own->sprite_winding - 1 : own->winding;   // a stroked path takes its sprite with EvenOdd (render.ml:1018)
      if (c.filter_kind == COH_FILTER_BLUR) {
        f.kernel_kind = c.filter_kernel & 255; f.r = c.filter_kernel >> 8;
        if ((f.kernel_kind != COH_CONV_UNIT && f.kernel_kind != COH_CONV_GAUSSIAN) || f.r <= 0 || f.r > 64) FAIL("Convolve.mkunit / mkxy: bad kernel");
      }
      if (c.filter_kind == COH_FILTER_MINUS) {   // filters.ml:295: hd scene
        f.head_abi = gj + 1;
        if (gj + 1 >= n_objs - n_background || objs[gj + 1].kind == COH_OBJ_GROUP_END || objs[gj + 1].kind == COH_OBJ_FILTER ||
            (objs[gj + 1].kind == COH_OBJ_GROUP_BEGIN && objs[gj + 1].filter_kind == COH_FILTER_READING_SCENE))
          FAIL("Filters.minus: no object below the filter (hd)");
      }
      if (geom_next) {
        // the geometry object as a scene of its own: its shape is the filter's, the alpha of its sprite the matte
        coh_scene_t sub = 0;
        if (coh_scene_create(ctx, objs + i + 1, gj - i, 0, edges, n_edges, points, n_points, &sub)) return 1;
        DevScene* ss = (DevScene*)sub;
        int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
        for (int li : ss->h_leaves) { const ObjRec& m = ss->h_objs[li]; x0 = std::min(x0, m.bx0); y0 = std::min(y0, m.by0); x1 = std::max(x1, m.bx1); y1 = std::max(y1, m.by1); }
        const int skip = gj;
        if (x0 > x1) { coh_scene_free(ctx, sub); i = skip; continue; }   // NullShape geometry
        f.bx0 = x0; f.by0 = y0; f.bx1 = x1; f.by1 = y1;
        f.gcx0 = floordiv(x0, 32) * 32; f.gcy0 = y0; f.gcnw = (x1 - f.gcx0) / 32 + 1; f.gch = y1 - y0 + 1;
        if (f.gch > 65535 || f.gcnw > 32767) { coh_scene_free(ctx, sub); FAIL("scene: filter geometry too large"); }
        if (subscene_translate(ctx, ss, objs + i + 1, gj - i, -f.gcx0, -f.gcy0)) { coh_scene_free(ctx, sub); return 1; }   // canvas coordinates from here on
        if (ss->n_leaves == 1 && ss->h_objs[ss->h_leaves[0]].kind == K_CONV) {
          // Convolved (k, g) as the geometry itself: sprite_of_basicshape convolves everywhere (render.ml:1023-1052), it
          // does not fill the eroded minshape like spriteof (1201-1204) — and the convolution of a raster whose interior
          // samples are not all 255 is not the fill there.  The canvas holds the convolution of the whole box: read all of it.
          const ObjRec& co = ss->h_objs[ss->h_leaves[0]];
          const size_t cw = (size_t)co.cv_nw * co.cv_h;
          CK(cudaMemsetAsync(ss->conv_bits + co.cv_bits + cw, 0, 4 * cw, ctx->stream));
        }
        f.geom_sub = ss; f.colour = 0xFFFFFFFFu;
        f.dx = c.dx; f.dy = c.dy; f.bx0 += c.dx; f.bx1 += c.dx; f.by0 += c.dy; f.by1 += c.dy;
        filters.push_back(f); filter_read_abi.push_back(c.filter_kind == COH_FILTER_SCENE ? c.first2 : -1);
        i = skip;
        continue;
      }
      if (own->count == 0) { i = gj; continue; }  // NullShape geometry: the filter touches nothing
      if (smear) {
        uint64_t rbits; memcpy(&rbits, &c.brush_radius, sizeof rbits);
        const std::pair<uint64_t, int> key(rbits, (int)(c.brush_opacity * 255.));
        auto it = stamp_cache.find(key);
        if (it == stamp_cache.end()) {
          f.stamp_off = (int)stamps.size();
          brush_stamp(c.brush_radius, c.brush_opacity, stamps, f.brush_r, &stamp_gauss);
          stamp_cache[key] = std::make_pair(f.stamp_off, f.brush_r);
        } else { f.stamp_off = it->second.first; f.brush_r = it->second.second; }
        f.first2 = c.first2; f.count2 = c.count2; f.colour = 0xFFFFFFFFu;
        int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
        for (int k = 0; k < c.count; k++) {
          const int32_t* p = points + 2 * ((size_t)c.first + k);
          x0 = std::min(x0, p[0]); x1 = std::max(x1, p[0]); y0 = std::min(y0, p[1]); y1 = std::max(y1, p[1]);
        }
        f.bx0 = x0 - f.brush_r; f.bx1 = x1 + f.brush_r; f.by0 = y0 - f.brush_r; f.by1 = y1 + f.brush_r;
      } else {
        EdgeBox eb = edge_bounds(edges + 4 * (size_t)own->first, own->count);
        shape_pixel_box(eb, f.bx0, f.by0, f.bx1, f.by1);
      }
      f.dx = c.dx; f.dy = c.dy;
      if (own != &c) { f.dx += own->dx; f.dy += own->dy; }
      f.bx0 += f.dx; f.bx1 += f.dx; f.by0 += f.dy; f.by1 += f.dy;
      filters.push_back(f); filter_read_abi.push_back(c.filter_kind == COH_FILTER_SCENE ? c.first2 : -1);
      i = gj;
      continue;
    }
    ObjRec o; memset(&o, 0, sizeof o);
    o.pretrans = c.pretrans; o.dx = c.dx; o.dy = c.dy;
    if (c.pretrans < -1 || c.pretrans > 255) FAIL("scene: pretrans out of range");
    o.depth = (int)eff_open.size();
    const int nesting = (int)open.size();   // the reference's nesting (group shapes, direct members), o.depth: what the walker sees
    if (o.depth > MAX_DEPTH) FAIL("scene: groups nested too deeply (MAX_DEPTH)");
    for (int d = 0; d < o.depth; d++) o.anc[d] = eff_open[d];
    o.fill.kind = c.fill_kind; o.fill.c0 = c.colour0; o.fill.c1 = c.colour1; o.fill.flags = c.fill_flags;
    for (int k = 0; k < 6; k++) o.fill.p[k] = c.fparam[k];
    switch (c.kind) {
      case COH_OBJ_GROUP_BEGIN: {
        // A Group with a filter object among its members (render.ml:988-1001: the members are rendered as a scene of
        // their own, so the filter's "objects below" are the rest of THIS list) is realised like a Convolved group with
        // no kernel: its members become a scene of their own, rendered once into the group's canvas.
        bool holds_filter = false;
        if (!c.convolve && c.filter_kind != COH_FILTER_READING_SCENE)
          for (int j = i + 1, nest = 1; j < n_objs && nest > 0; j++) {
            if (objs[j].kind == COH_OBJ_GROUP_BEGIN) nest++;
            else if (objs[j].kind == COH_OBJ_GROUP_END) nest--;
            else if (objs[j].kind == COH_OBJ_FILTER) { holds_filter = true; break; }
          }
        // ... and so is a group that would be nested more deeply than the walker's accumulator stack (MAX_DEPTH levels,
        // counted after first-member groups have dissolved): the reference has no nesting limit
        const bool too_deep = !c.convolve && c.filter_kind != COH_FILTER_READING_SCENE && o.depth >= MAX_DEPTH && !(o.pretrans < 0 && n_children.back() == 0);
        if (c.convolve || holds_filter || too_deep) {
          // Convolved (kernel, Group members) (render.ml:63, 1023-1052 with a Group child; shapes render.ml:536-555, where
          // findfill of a Group is "fancy": minshape null).  The members become a scene of their own, rendered once into
          // the object's canvas and convolved there (below); the object itself is one leaf, like Convolved (Basic Path).
          const int ck = c.convolve & 255, cr = c.convolve >> 8;
          if (c.convolve && ((ck != COH_CONV_UNIT && ck != COH_CONV_GAUSSIAN) || cr <= 0 || cr > 64)) FAIL("Convolve.mkunit / mkxy: bad kernel");
          if (c.filter_kind == COH_FILTER_READING_SCENE) FAIL("scene: a reading-scene group cannot be Convolved");
          int j = i + 1, nest = 1;
          for (; j < n_objs; j++) {
            const coh_object& m = objs[j];
            if (j == n_objs - n_background) break;   // (a group may not straddle the two lists)
            if (m.kind == COH_OBJ_GROUP_BEGIN) nest++;
            else if (m.kind == COH_OBJ_GROUP_END) { if (--nest == 0) break; }
            else if (m.kind == COH_OBJ_FILTER) {
              if (m.filter_kind == COH_FILTER_SCENE) FAIL("scene: a filter with a caller-built reading scene must be a top-level member of the scene list");
            } else if (m.kind != COH_OBJ_PRIMITIVE && m.fill_kind != COH_FILL_PLAIN)
              // the span-start fill quirk (polygon.ml:736) makes a fancy-filled member depend on the region requested at
              // render time, which a canvas rendered once cannot follow
              FAIL("scene: Convolved groups (and groups holding filters) with fancy-filled members are not supported yet");
          }
          if (j >= n_objs || nest != 0) FAIL("scene: unterminated group");
          if (j == i + 1) FAIL("Empty groups aren't allowed");   // render.ml:317
          coh_scene_t sub = 0;
          if (coh_scene_create(ctx, objs + i + 1, j - i - 1, 0, edges, n_edges, points, n_points, &sub)) return 1;
          DevScene* ss = (DevScene*)sub;
          conv_groups.push_back({(int)recs.size(), ck, cr, ss, objs + i + 1, j - i - 1});
          skip_to = j;   // the members belong to the Convolved object
          int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
          for (int li : ss->h_leaves) { const ObjRec& m = ss->h_objs[li]; x0 = std::min(x0, m.bx0); y0 = std::min(y0, m.by0); x1 = std::max(x1, m.bx1); y1 = std::max(y1, m.by1); }
          for (const DevScene::FilterRec& m : ss->filters) { x0 = std::min(x0, m.bx0); y0 = std::min(y0, m.by0); x1 = std::max(x1, m.bx1); y1 = std::max(y1, m.by1); }   // a filter's shape is its geometry's (render.ml:472-474)
          if (x0 > x1) { rec_of_abi[i] = -1; i = j; conv_groups.back().rec = -1; continue; }   // nothing to draw
          o.kind = K_CONV; o.fill.kind = 0; o.fill.c0 = 0;
          o.bx0 = x0; o.by0 = y0; o.bx1 = x1; o.by1 = y1;
          o.cv_x0 = floordiv(o.bx0 - 2 * cr, 32) * 32; o.cv_y0 = o.by0 - 2 * cr;
          o.cv_nw = (o.bx1 + 2 * cr - o.cv_x0) / 32 + 1; o.cv_h = o.by1 + 2 * cr - o.cv_y0 + 1;
          o.bx0 -= cr; o.bx1 += cr; o.by0 -= cr; o.by1 += cr;
          o.cv_bits = (int)conv_words; conv_words += 2 * (size_t)o.cv_nw * o.cv_h;
          o.cv_px = (int)conv_pixels; conv_pixels += (size_t)o.cv_nw * 32 * o.cv_h;
          if (conv_words > 0x7FFFFFF0ull || conv_pixels > 0x7FFFFFF0ull || o.cv_h > 65535 || o.cv_nw > 32767) FAIL("scene: Convolved canvases too large");
          break;
        }
        }
        o.kind = K_GROUP;
        if (o.depth >= MAX_DEPTH) FAIL("scene: groups nested too deeply (MAX_DEPTH)");
        recs.push_back(o); open.push_back((int)recs.size() - 1); open_reading.push_back(-1);
        real_depth.resize(recs.size(), 0); real_depth.back() = nesting;
        {
          const bool flat = o.pretrans < 0 && n_children.back() == 0;
          n_children.back()++;
          open_flat.push_back(flat ? 1 : 0); n_children.push_back(0);
          if (!flat) eff_open.push_back((int)recs.size() - 1);
        }
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
        const EdgeBox eb = pre_bounds[i].a;
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
        edge_own.add(c.first, c.count, (int)recs.size());
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
          const EdgeBox eb = side ? pre_bounds[i].b : pre_bounds[i].a;
          int x0, y0, x1, y1;
          shape_pixel_box(eb, x0, y0, x1, y1);
          o.bx0 = std::min(o.bx0, x0); o.by0 = std::min(o.by0, y0); o.bx1 = std::max(o.bx1, x1); o.by1 = std::max(o.by1, y1);
          const int r0 = floordiv(eb.ymin - 16 + 31, 32), r1 = floordiv(eb.ymax + 67, 32);
          if (total_rows + (r1 - r0 + 1) > 0x7FFFFFF0LL) FAIL("scene: too many object rows for the row-edge table");
          if (side) { o.b_ry0 = r0; o.b_ry1 = r1; o.b_row_base = (int)total_rows; } else { o.ry0 = r0; o.ry1 = r1; o.row_base = (int)total_rows; }
          total_rows += r1 - r0 + 1;
          edge_own.add(f, n, (int)recs.size());
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
        if (c.winding != COH_BRUSH_GAUSSIAN && c.winding != COH_BRUSH_DUMMY) FAIL("scene: bad brush kind");
        o.kind = K_BRUSH; o.first = c.first; o.count = c.count;
        if (c.count == 0) continue;
        if (c.winding == COH_BRUSH_DUMMY) {
          // Dummy (r, r) (brush.ml:14-22, 178-181): every pixel of the stroke's shape in opaque white — a stamp of 255s
          if (c.brush_radius > 4096. || c.brush_radius != (double)(int)c.brush_radius) FAIL("scene: the radius of a dummy brush is a small integer");
          o.fill.kind = 0; o.fill.c0 = 0xFFFFFFFFu;
          const std::pair<uint64_t, int> key((uint64_t)c.brush_radius, -1);
          auto it = stamp_cache.find(key);
          if (it == stamp_cache.end()) {
            o.stamp_off = (int)stamps.size(); o.brush_r = (int)c.brush_radius;
            stamps.resize(stamps.size() + (size_t)(2 * o.brush_r + 1) * (2 * o.brush_r + 1), (uint8_t)255);
            stamp_cache[key] = std::make_pair(o.stamp_off, o.brush_r);
          } else { o.stamp_off = it->second.first; o.brush_r = it->second.second; }
        } else {
          // the stamp is a function of (radius, toint (opacity *. 255.)) alone (brush.ml:60-92): strokes share it
          // (10^4 strokes of a scene use a few hundred distinct stamps; the exponentials are the cost of this loop)
          uint64_t rbits; memcpy(&rbits, &c.brush_radius, sizeof rbits);
          const std::pair<uint64_t, int> key(rbits, (int)(c.brush_opacity * 255.));
          auto it = stamp_cache.find(key);
          if (it == stamp_cache.end()) {
            o.stamp_off = (int)stamps.size();
            brush_stamp(c.brush_radius, c.brush_opacity, stamps, o.brush_r, &stamp_gauss);
            stamp_cache[key] = std::make_pair(o.stamp_off, o.brush_r);
          } else { o.stamp_off = it->second.first; o.brush_r = it->second.second; }
        }
        const int x0 = pre_bounds[i].a.xmin, x1 = pre_bounds[i].a.xmax, y0 = pre_bounds[i].a.ymin, y1 = pre_bounds[i].a.ymax;
        o.bx0 = x0 - o.brush_r; o.bx1 = x1 + o.brush_r; o.by0 = y0 - o.brush_r; o.by1 = y1 + o.brush_r;
        o.ry0 = o.by0; o.ry1 = o.by1;   // object-frame rows (the alias offset is added to the box below)
        o.bc_x0 = floordiv(o.bx0, 32); o.bc_y0 = floordiv(o.by0, CELL_H);
        o.bc_nx = floordiv(o.bx1, 32) - o.bc_x0 + 1; o.bc_ny = floordiv(o.by1, CELL_H) - o.bc_y0 + 1;
        if (total_brush_rows + (long long)o.bc_nx * o.bc_ny > 0x7FFFFFF0LL) FAIL("scene: too many brush cells");
        o.bc_base = (int)total_brush_rows; total_brush_rows += (long long)o.bc_nx * o.bc_ny;
        point_own.add(c.first, c.count, (int)recs.size());
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
    real_depth.resize(recs.size(), 0); real_depth.back() = nesting;
    n_children.back()++;
    rec_of_abi[i] = (int)recs.size() - 1;
    ids.resize(recs.size(), -1); ids.back() = c.id;
    leaves.push_back((int)recs.size() - 1);
    if (skip_to >= 0) i = skip_to;
  }
  mark("object loop");
  if (edge_own.overlapping()) FAIL("scene: objects may not share edges");
  if (point_own.overlapping()) FAIL("scene: objects may not share brush points");
  if (open.size() != 1 || cur_reading >= 0) FAIL("scene: unterminated group");
  if (n_scene_leaves < 0) n_scene_leaves = (int)leaves.size();
  if (n_front_leaves < 0) n_front_leaves = (int)leaves.size();
  for (DevScene::FilterRec& f : filters) {
    if (f.kind != COH_FILTER_MINUS) continue;
    // the list continues after the head object's last leaf (records and leaves are made in list order)
    const int hr = rec_of_abi[f.head_abi];
    f.head_l1 = f.pos;
    if (hr >= 0) {
      group_last.resize(recs.size(), -1);
      const int last = recs[hr].kind == K_GROUP ? group_last[hr] : hr;
      const int end = n_scene_leaves;
      while (f.head_l1 < end && leaves[f.head_l1] <= last) f.head_l1++;
    }
  }
  for (size_t k = 0; k < filters.size(); k++) {
    if (filter_read_abi[k] < 0) continue;
    auto it = reading.find(filter_read_abi[k]);
    if (it == reading.end()) FAIL("scene: filter without its reading-scene group");
    filters[k].read0 = it->second.first; filters[k].read1 = it->second.second;
  }
  group_last.resize(recs.size(), -1);
  ids.resize(recs.size(), -1);
  real_depth.resize(recs.size(), 0);
  // Partial-sprite cache (render.ml:1169-1242): a top-level Group of the scene list that carries an id (the only
  // kind of object whose sprite the reference ever finds again: members get fresh ids on every render, render.ml:993)
  // is given a sprite leaf — shape planes and an RGBA8 canvas in its own frame, laid out like a Convolved object's —
  // and a second leaf list in which that leaf stands for all its members.
  std::vector<SpriteEntry> sprites;
  const int n_real_recs = (int)recs.size();
  if (filters.empty())
    for (int g = 0; g < n_real_recs; g++) {
      if (recs[g].kind != K_GROUP || real_depth[g] != 1 || ids[g] < 0 || recs[g].pretrans >= 0) continue;
      if (!(recs[recs[g].anc[0]].flags & OF_ROOT_SCENE)) continue;
      int l0 = -1, l1 = -1;
      for (int li = 0; li < (int)leaves.size(); li++)
        if (leaves[li] > g && leaves[li] <= group_last[g]) { if (l0 < 0) l0 = li; l1 = li + 1; }
      if (l0 < 0 || l1 - l0 < 2) continue;
      bool ok = true;
      int bx0 = INT32_MAX, by0 = INT32_MAX, bx1 = INT32_MIN, by1 = INT32_MIN;
      for (int li = l0; li < l1 && ok; li++) {
        const ObjRec& m = recs[leaves[li]];
        ok = (m.kind == K_PATH || m.kind == K_PRIM) && m.fill.kind == 0 && m.dx == 0 && m.dy == 0;
        bx0 = std::min(bx0, m.bx0); by0 = std::min(by0, m.by0); bx1 = std::max(bx1, m.bx1); by1 = std::max(by1, m.by1);
      }
      if (!ok) continue;
      ObjRec o; memset(&o, 0, sizeof o);
      o.kind = K_CONV; o.pretrans = -1; o.depth = 1; o.anc[0] = recs[g].anc[0];
      o.bx0 = bx0; o.by0 = by0; o.bx1 = bx1; o.by1 = by1;
      o.cv_x0 = floordiv(bx0, 32) * 32; o.cv_y0 = by0; o.cv_nw = (bx1 - o.cv_x0) / 32 + 1; o.cv_h = by1 - by0 + 1;
      const size_t words = (size_t)o.cv_nw * o.cv_h, bytes = words * 4 * 3 + words * 32 * 4;   // S, M, V planes + canvas
      if (bytes > ctx->cache_max / 2) continue;                      // cache.ml: an item larger than half the cache is not kept
      if (conv_words + 2 * words > 0x7FFFFFF0ull || conv_pixels + words * 32 > 0x7FFFFFF0ull) continue;
      o.cv_bits = (int)conv_words; conv_words += 2 * words;
      o.cv_px = (int)conv_pixels; conv_pixels += words * 32;
      SpriteEntry e; e.grp = g; e.l0 = l0; e.l1 = l1; e.leaf_rec = (int)recs.size(); e.plane_words = words; e.bytes = bytes;
      recs.push_back(o);
      sprites.push_back(e);
    }
  group_last.resize(recs.size(), -1);
  ids.resize(recs.size(), -1);
  real_depth.resize(recs.size(), 0);
  mark("sprite entries");
  std::vector<int> filter_taps;
  for (DevScene::FilterRec& f : filters) {
    f.taps_off = (int)filter_taps.size(); f.taps_total = 0;
    if (f.kind != COH_FILTER_BLUR || f.kernel_kind != COH_CONV_GAUSSIAN) continue;
    for (int i = -f.r; i <= f.r; i++) {   // Convolve.mkgaussian r (convolve.ml:60-70)
      double xr = (double)i / (double)f.r, yr = 0. / (double)f.r;
      double gg = exp(-(xr * xr + yr * yr)) / 2.;
      int v = (int)((double)(4 * f.r * f.r) * gg + 0.5);
      filter_taps.push_back(v); f.taps_total += v;
    }
  }
  if (!filter_taps.empty()) {
    CK(DMALLOC(&s->filter_taps, sizeof(int) * filter_taps.size()));
    CK(cudaMemcpyAsync(s->filter_taps, filter_taps.data(), sizeof(int) * filter_taps.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  s->filters = filters; s->n_scene_leaves = n_scene_leaves; s->n_front_leaves = n_front_leaves;
  s->n_objs = (int)recs.size(); s->n_leaves = (int)leaves.size(); s->n_edges = n_edges; s->n_points = n_points;
  s->sprites = sprites;
  if (!sprites.empty()) {   // the collapsed leaf list
    size_t k = 0;
    for (int li = 0; li < (int)leaves.size(); li++) {
      while (k < sprites.size() && li >= sprites[k].l1) k++;
      if (k < sprites.size() && li >= sprites[k].l0) { if (li == sprites[k].l0) s->sp.h_leaves.push_back(sprites[k].leaf_rec); continue; }
      s->sp.h_leaves.push_back(leaves[li]);
    }
    s->sp.n = (int)s->sp.h_leaves.size();
  }
  s->rec_of_abi = rec_of_abi; s->group_last = group_last; s->ids = ids; s->real_depth = real_depth;
  s->group_off.assign(recs.size(), make_int2(0, 0));
  for (int ri = 0; ri < n_real_recs; ri++) {
    const ObjRec& o = recs[ri];
    if (o.kind != K_GROUP && o.kind != K_PRIM && o.fill.kind != 0) s->has_fancy = true;
    if (o.kind == K_BRUSH || o.kind == K_CONV) s->extras = std::max(s->extras, 1);
    if (o.kind == K_CPG || !filters.empty()) s->extras = 2;  // the filter passes need the walker variant that can continue a frame
  }
  // Flat scenes — every leaf a direct member of one of the two root lists, plain-filled paths and primitives only —
  // take the row compositor (k_comp_rows) in three-phase frames.  The background list is composited as a list of
  // its own (render.ml:1364-1365); walking it under the scene's accumulator gives the same pixels when it has one
  // member, or only opaque primitives (the first one covering a pixel finishes it either way).
  std::vector<int2> attr(recs.size(), make_int2(0, 0));   // (uploaded asynchronously: lives until the synchronisation below)
  {
    bool flat = true;
    int n_bg = 0; bool bg_opaque_prims = true;
    for (int li : leaves) {
      const ObjRec& o = recs[li];
      flat = flat && o.depth == 1 && (o.kind == K_PATH || o.kind == K_PRIM || o.kind == K_CONV) && o.fill.kind == 0;
      const bool bg = (recs[o.anc[0]].flags & OF_ROOT_BACKGROUND) != 0;
      if (bg) { n_bg++; bg_opaque_prims = bg_opaque_prims && o.kind == K_PRIM && (o.fill.c0 >> 24) == 255u && o.pretrans < 0; }
      attr[li] = make_int2((int)o.fill.c0, (o.kind == K_PATH ? 1 : 0) | (bg ? 2 : 0) | ((o.flags & OF_OCCLUDES) ? 4 : 0) | (o.kind == K_CONV ? 8 : 0) | ((o.pretrans + 1) << 8));
    }
    s->flat_ok = flat && (n_bg <= 1 || bg_opaque_prims);
    CK(DMALLOC(&s->attr, sizeof(int2) * recs.size()));
    CK(cudaMemcpyAsync(s->attr, attr.data(), sizeof(int2) * recs.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(DMALLOC(&s->objs, sizeof(ObjRec) * recs.size()));
  CK(cudaMemcpyAsync(s->objs, recs.data(), sizeof(ObjRec) * recs.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(DMALLOC(&s->leaves, sizeof(int) * std::max<size_t>(leaves.size(), 1)));
  if (!leaves.empty()) CK(cudaMemcpyAsync(s->leaves, leaves.data(), sizeof(int) * leaves.size(), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<int4> boxes(leaves.size());
  for (size_t i = 0; i < leaves.size(); i++) { const ObjRec& o = recs[leaves[i]]; boxes[i] = make_int4(o.bx0, o.by0, o.bx1, o.by1); }
  CK(DMALLOC(&s->leaf_box, sizeof(int4) * std::max<size_t>(leaves.size(), 1)));
  if (!leaves.empty()) CK(cudaMemcpyAsync(s->leaf_box, boxes.data(), sizeof(int4) * boxes.size(), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<int4> sp_boxes(s->sp.h_leaves.size());
  if (s->sp.n > 0) {
    for (size_t i = 0; i < sp_boxes.size(); i++) { const ObjRec& o = recs[s->sp.h_leaves[i]]; sp_boxes[i] = make_int4(o.bx0, o.by0, o.bx1, o.by1); }
    CK(DMALLOC(&s->sp.leaves, sizeof(int) * s->sp.n));
    CK(DMALLOC(&s->sp.leaf_box, sizeof(int4) * s->sp.n));
    CK(cudaMemcpyAsync(s->sp.leaves, s->sp.h_leaves.data(), sizeof(int) * s->sp.n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(s->sp.leaf_box, sp_boxes.data(), sizeof(int4) * s->sp.n, cudaMemcpyHostToDevice, ctx->stream));
  }
  mark("records, leaves, boxes");
  if (upload_edges(ctx, edges, n_edges, &s->edges)) return 1;
  if (up.th.joinable()) { up.th.join(); CK(up.err); up.taken = true; }   // uploaded meanwhile
  else {
    CK(DMALLOC(&s->points, sizeof(int2) * std::max(n_points, 1)));
    if (n_points > 0) CK(cudaMemcpyAsync(s->points, points, sizeof(int2) * n_points, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(DMALLOC(&s->stamps, std::max<size_t>(stamps.size(), 1)));
  if (!stamps.empty()) CK(cudaMemcpyAsync(s->stamps, stamps.data(), stamps.size(), cudaMemcpyHostToDevice, ctx->stream));
  mark("edges, points, stamps");
  // K1 edge binning: count -> scan -> fill
  {
    int* d_edge_obj = nullptr; int* d_counts = nullptr;
    size_t slots = (size_t)std::max<long long>(total_rows, 1);
    if (expand_owners(ctx, edge_own, n_edges, &d_edge_obj)) return 1;
    CK(DMALLOC(&d_counts, sizeof(int) * slots));
    CK(DMALLOC(&s->rowedge_ptr, sizeof(int) * (slots + 1)));
    CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * slots, ctx->stream));
    if (n_edges > 0) { k_rowedges<false><<<cdiv(n_edges * 32, 256), 256, 0, ctx->stream>>>(s->edges, d_edge_obj, n_edges, s->objs, d_counts, nullptr, nullptr); LAUNCHED(); }
    if (exclusive_scan(ctx, d_counts, s->rowedge_ptr, (int)slots, nullptr)) return 1;
    // size of the lists: the same row range per edge as k_rowedges, summed on the host (no device round trip:
    // a device-to-host read here would queue behind an asynchronous framebuffer read-back of the previous frame)
    long long total = 0;
    for (const int4& rg : edge_own.ranges)
     for (int e = rg.x; e < rg.x + rg.y; e++) {
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
  mark("row-edge lists");
  // Convolved objects (render.ml:1023-1052): AA-rasterise the whole (twice bloated) box of the child,
  // X pass, Y pass; keep the shape / minshape bit-rows and the convolved canvas resident.
  if (conv_words > 0) {
    CK(DMALLOC(&s->conv_bits, sizeof(uint32_t) * conv_words));
    CK(DMALLOC(&s->conv_px, sizeof(uint32_t) * conv_pixels));
  }
  // sprite leaves: shape plane = union of the members' shapes (render.ml:476-496), minshape plane null, pshape empty
  for (SpriteEntry& e : s->sprites) {
    const ObjRec& o = recs[e.leaf_rec];
    const int nw = o.cv_nw, h = o.cv_h;
    uint32_t* S = s->conv_bits + o.cv_bits;
    uint32_t *Cs = nullptr;
    CK(DMALLOC(&Cs, 4 * e.plane_words));
    CK(cudaMemsetAsync(S, 0, 4 * 2 * e.plane_words, ctx->stream));
    CK(DMALLOC(&e.valid, 4 * e.plane_words));
    CK(cudaMemsetAsync(e.valid, 0, 4 * e.plane_words, ctx->stream));
    for (int li = e.l0; li < e.l1; li++) {
      const ObjRec& m = recs[leaves[li]];
      const int r0 = m.by0 - o.cv_y0, rows = m.by1 - m.by0 + 1;   // the member's rows of the planes
      if (m.kind == K_PATH) {
        k_scan_rows<<<dim3(cdiv(rows, 64), cdiv(nw, SCAN_CHUNK_WORDS)), 64, 0, ctx->stream>>>(s->edges + m.first, m.count, m.winding, m.by0, rows, o.cv_x0, nw,
                                                                                                 S + (size_t)r0 * nw, Cs + (size_t)r0 * nw, ctx->d_error); LAUNCHED();
      } else {
        k_fill_box_bits<<<dim3(cdiv(nw, 128), rows), 128, 0, ctx->stream>>>(Cs + (size_t)r0 * nw, rows, nw, o.cv_x0, m.by0, m.prim[0], m.prim[1], m.prim[2], m.prim[3]); LAUNCHED();
        k_bitop<<<(unsigned)(((size_t)rows * nw + 255) / 256), 256, 0, ctx->stream>>>(S + (size_t)r0 * nw, Cs + (size_t)r0 * nw, S + (size_t)r0 * nw, (size_t)rows * nw, 0); LAUNCHED();
      }
    }
    DFREE(Cs);
    CK(DMALLOC(&e.d_missing, sizeof(int)));
    CK(cudaMallocHost(&e.h_missing, sizeof(int)));
    CK(cudaEventCreateWithFlags(&e.ev, cudaEventDisableTiming));
  }
  if (!conv_list.empty()) {
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
  for (ConvGroup& cg : conv_groups)
    if (cg.rec >= 0 && realize_convolved_group(ctx, s, recs[cg.rec], cg)) return 1;
  mark("sprites, Convolved");
  if (total_brush_rows > 0) {
    int* d_point_obj = nullptr;
    if (expand_owners(ctx, point_own, n_points, &d_point_obj)) return 1;
    CK(DMALLOC(&s->brush_ranges, sizeof(int2) * (size_t)total_brush_rows));
    k_fill_int2<<<cdiv((int)total_brush_rows, 256), 256, 0, ctx->stream>>>(s->brush_ranges, (int)total_brush_rows, make_int2(INT32_MAX, -1)); LAUNCHED();
    k_brush_cells<<<cdiv(n_points, 256), 256, 0, ctx->stream>>>(s->points, d_point_obj, n_points, s->objs, s->brush_ranges); LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));
    DFREE(d_point_obj);
  }
  mark("brush cells");
  *out = (coh_scene_t)holder.release();
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
  // the pair list of three-phase frames packs (tile column << 16 | row) into one word
  if (height > 65535 || width > 32 * 32767) FAIL("coh_fb_configure: framebuffer larger than 1048544 x 65535 pixels");
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
