// host_polygon.inl — part of coherence_b200.cu (one translation unit; included in order): Polygon entry points (scan conversion, AA opacity, sprites) and Convolve.convolve_sprite.

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
    if (*ctx->h_error == 5) { ctx->err = std::string(what) + ": coh_frame_wait gave up after 4 s (a peer never signalled its frame)"; return 1; }
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

// ---------------------------------------------------------------------------------------
// N2: Polygon.edgelist_of_path on the device.  The flattened edges stay in HBM; `box` (host, 4 ints:
// sub-bin x min / y min / x max / y max) and the edge count come back in one small read.
// ---------------------------------------------------------------------------------------
static int flatten_on_device(coh_ctx* ctx, const double* segs, int n_segs, int4** d_edges, int* n_edges, int box[4]) {
  *d_edges = nullptr; *n_edges = 0;
  if (n_segs <= 0) return 0;
  double* d_segs = nullptr; int *counts = nullptr, *offs = nullptr, *d_box = nullptr;
  CK(DMALLOC(&d_segs, sizeof(double) * 9 * (size_t)n_segs));
  CK(DMALLOC(&counts, sizeof(int) * n_segs)); CK(DMALLOC(&offs, sizeof(int) * ((size_t)n_segs + 1))); CK(DMALLOC(&d_box, 4 * sizeof(int)));
  CK(cudaMemcpyAsync(d_segs, segs, sizeof(double) * 9 * (size_t)n_segs, cudaMemcpyHostToDevice, ctx->stream));
  const int init[4] = {INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN};
  CK(cudaMemcpyAsync(d_box, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  k_flatten<false><<<cdiv(n_segs, 128), 128, 0, ctx->stream>>>(d_segs, n_segs, counts, nullptr, nullptr, nullptr, ctx->d_error); LAUNCHED();
  if (exclusive_scan(ctx, counts, offs, n_segs, nullptr)) return 1;
  int total = 0;
  CK(cudaMemcpyAsync(&total, offs + n_segs, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(DMALLOC(d_edges, sizeof(int4) * (size_t)std::max(total, 1)));
  k_flatten<true><<<cdiv(n_segs, 128), 128, 0, ctx->stream>>>(d_segs, n_segs, nullptr, offs, *d_edges, d_box, ctx->d_error); LAUNCHED();
  CK(cudaMemcpyAsync(box, d_box, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_error, ctx->d_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(d_segs); DFREE(counts); DFREE(offs); DFREE(d_box);
  if (*ctx->h_error) {
    cudaMemsetAsync(ctx->d_error, 0, sizeof(int), ctx->stream);
    DFREE(*d_edges);
    FAIL("coh_edgelist_of_path: a curve needs more than 40 levels of subdivision");
  }
  *n_edges = total;
  return 0;
}
int coh_edgelist_of_path(coh_ctx* ctx, const double* segs, int32_t n_segs, int32_t* edges_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  int4* d_edges = nullptr; int n = 0, box[4];
  if (flatten_on_device(ctx, segs, n_segs, &d_edges, &n, box)) return 1;
  *n_out = n;
  if (n > 0 && cap > 0) {
    CK(cudaMemcpyAsync(edges_out, d_edges, sizeof(int4) * (size_t)std::min<int64_t>(n, cap), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  DFREE(d_edges);
  return 0;
}
int coh_shapeminshape_of_path(coh_ctx* ctx, const double* segs, int32_t n_segs, int32_t winding, coh_shape_t* shape, coh_shape_t* minshape) {
  CK(cudaSetDevice(ctx->device));
  *shape = 0; *minshape = 0;
  if (winding != COH_NONZERO && winding != COH_EVENODD) FAIL("bad winding rule");
  int4* raw = nullptr; int n = 0, box[4];
  if (flatten_on_device(ctx, segs, n_segs, &raw, &n, box)) return 1;
  if (n <= 0) { DFREE(raw); return 0; }
  EdgeRec* d_edges = nullptr;
  CK(DMALLOC(&d_edges, sizeof(EdgeRec) * (size_t)n));
  k_prep_edges<<<cdiv(n, 256), 256, 0, ctx->stream>>>(raw, d_edges, n); LAUNCHED();
  int px0, py0, px1, py1; shape_pixel_box(EdgeBox{box[0], box[2], box[1], box[3]}, px0, py0, px1, py1);
  int rc = shapes_from_device_edges(ctx, d_edges, n, winding, px0, py0, px1, py1, shape, minshape, "coh_shapeminshape_of_path");
  DFREE(raw); DFREE(d_edges);
  return rc;
}

// Shapes.strokepath (shapes.ml:529-530): the outline from the host stroker (host_stroke.cpp), flattened by k_flatten
static int stroke_outline(coh_ctx* ctx, const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths,
                          std::vector<double>& outline, int32_t* winding) {
  if (!spec || n_subpaths < 0) FAIL("strokepath: bad arguments");
  if (spec->startcap < 0 || spec->startcap > 2 || spec->endcap < 0 || spec->endcap > 2 || spec->join < 0 || spec->join > 2) FAIL("strokepath: bad cap or join");
  std::vector<int32_t> counts((size_t)std::max(n_subpaths, 1));
  int32_t m = 0;
  int64_t n = coh_host_strokepath(spec, segs, subpath_segs, n_subpaths, nullptr, 0, counts.data(), 0, &m, winding);
  if (n < 0) FAIL("Shapes.joinsegments: a rail that ends in a curve cannot be joined");
  outline.resize((size_t)std::max<int64_t>(n, 1) * 9);
  if (coh_host_strokepath(spec, segs, subpath_segs, n_subpaths, outline.data(), n, counts.data(), 0, &m, winding) != n) FAIL("strokepath: inconsistent outline");
  outline.resize((size_t)n * 9);
  return 0;
}
int coh_strokepath(coh_ctx* ctx, const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths,
                   int32_t* edges_out, int64_t cap, int64_t* n_out, int32_t* winding_out) {
  *n_out = 0;
  std::vector<double> outline; int32_t winding = COH_EVENODD;
  if (stroke_outline(ctx, spec, segs, subpath_segs, n_subpaths, outline, &winding)) return 1;
  if (winding_out) *winding_out = winding;
  CK(cudaSetDevice(ctx->device));
  int4* d_edges = nullptr; int n = 0, box[4];
  if (flatten_on_device(ctx, outline.data(), (int)(outline.size() / 9), &d_edges, &n, box)) return 1;
  *n_out = n;
  if (n > 0) {
    std::vector<int4> e((size_t)n);
    CK(cudaMemcpyAsync(e.data(), d_edges, sizeof(int4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::stable_sort(e.begin(), e.end(), [](const int4& a, const int4& b) { return std::max(a.y, a.w) > std::max(b.y, b.w); });   // polygon.ml:243-244
    if (cap > 0) memcpy(edges_out, e.data(), sizeof(int4) * (size_t)std::min<int64_t>(n, cap));
  }
  DFREE(d_edges);
  return 0;
}
int coh_shapeminshape_of_stroke(coh_ctx* ctx, const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths,
                                coh_shape_t* shape, coh_shape_t* minshape) {
  *shape = 0; *minshape = 0;
  std::vector<double> outline; int32_t winding = COH_EVENODD;
  if (stroke_outline(ctx, spec, segs, subpath_segs, n_subpaths, outline, &winding)) return 1;
  return coh_shapeminshape_of_path(ctx, outline.data(), (int32_t)(outline.size() / 9), winding, shape, minshape);
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
// Offset of every row's first pixel in canonical span order (device, n_rows + 1 ints): row pixel counts + scan.
static int shape_pixel_offsets(coh_ctx* ctx, const DevShape* s, int** d_off) {
  int* counts = nullptr;
  *d_off = nullptr;
  if (s->card > 0x7FFFFFF0LL) FAIL("span set too large for a per-pixel export");
  CK(DMALLOC(&counts, sizeof(int) * std::max(s->n_rows, 1)));
  CK(DMALLOC(d_off, sizeof(int) * (s->n_rows + 1)));
  k_row_pixels<<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, s->n_rows, counts); LAUNCHED();
  if (exclusive_scan(ctx, counts, *d_off, s->n_rows, nullptr)) return 1;
  DFREE(counts);
  return 0;
}
// AA opacity of every pixel of `s` in canonical span order, device resident (card bytes)
static int polygon_opacity_spans(coh_ctx* ctx, const int32_t* edges, int n_edges, int winding, const DevShape* s, uint8_t** d_op, int** d_off) {
  uint8_t* dense = nullptr; int wx0, nw;
  *d_op = nullptr;
  if (shape_pixel_offsets(ctx, s, d_off)) return 1;
  CK(DMALLOC(d_op, (size_t)std::max<long long>(s->card, 1)));
  if (n_edges <= 0) { CK(cudaMemsetAsync(*d_op, 0, (size_t)s->card, ctx->stream)); return 0; }  // empty scaled shape: coverage 0
  if (polygon_opacity_dense(ctx, edges, n_edges, winding, s, &dense, &wx0, &nw)) return 1;
  k_gather_spans<uint8_t><<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, *d_off, s->n_rows, wx0, nw * 32, dense, *d_op); LAUNCHED();
  DFREE(dense);
  return 0;
}
int coh_polygon_opacity(coh_ctx* ctx, const int32_t* edges, int32_t n_edges, int32_t winding, coh_shape_t shp,
                        uint8_t* out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  if (!shp) return 0;
  DevShape* s = (DevShape*)shp;
  if (s->card > cap) FAIL("coh_polygon_opacity: buffer too small");
  uint8_t* d_op = nullptr; int* d_off = nullptr;
  if (polygon_opacity_spans(ctx, edges, n_edges, winding, s, &d_op, &d_off)) return 1;
  CK(cudaMemcpyAsync(out, d_op, (size_t)s->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(d_op); DFREE(d_off);
  *n_out = s->card;
  return 0;
}
int coh_polygon_sprite(coh_ctx* ctx, const coh_object* fill, const int32_t* edges, int32_t n_edges, int32_t winding,
                       coh_shape_t shp, uint32_t* out, int64_t cap, int64_t* n_out) {
  // polygon.ml:729-746: per span, colour = dissolve (fillsingle x_spanstart y) opacity — the opacity from the
  // AA kernel, fill and dissolve by k_sprite_fill, all on the device; only the finished sprite crosses the ABI.
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  if (!shp) return 0;
  DevShape* s = (DevShape*)shp;
  if (s->card > cap) FAIL("coh_polygon_sprite: buffer too small");
  if (fill->fill_kind < COH_FILL_PLAIN || fill->fill_kind > COH_FILL_RADIAL) FAIL("coh_polygon_sprite: bad fill kind");
  uint8_t* d_op = nullptr; int* d_off = nullptr; uint32_t* d_out = nullptr;
  if (polygon_opacity_spans(ctx, edges, n_edges, winding, s, &d_op, &d_off)) return 1;
  FillRec f; f.kind = fill->fill_kind; f.c0 = fill->colour0; f.c1 = fill->colour1; f.flags = fill->fill_flags;
  for (int i = 0; i < 6; i++) f.p[i] = fill->fparam[i];
  CK(DMALLOC(&d_out, 4 * (size_t)std::max<long long>(s->card, 1)));
  k_sprite_fill<<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, d_off, s->y0, s->n_rows, f, d_op, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(out, d_out, 4 * (size_t)s->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(d_op); DFREE(d_off); DFREE(d_out);
  *n_out = s->card;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Brush strokes outside a scene (brush.mli:20-27): shape_of_brushstroke, sprite_of_brushstroke, smear.  The stroke
// crosses the boundary as its rounded stamp points (Brush.points_of_brushstroke, brush.ml:126-130, 172) and a BRUSH
// object record (brush_radius, brush_opacity, winding = COH_BRUSH_*, fill).
// ---------------------------------------------------------------------------------------
static FillRec fillrec_of(const coh_object* fill);
typedef std::map<uint64_t, std::vector<uint8_t>> StampGauss;
static void brush_stamp(double radius, double opacity, std::vector<uint8_t>& out, int& r_out, StampGauss* gauss);   // host_scene.inl
static int brush_radius_of(coh_ctx* ctx, const coh_object* b, int* r) {
  if (!(b->brush_radius >= 0. && b->brush_radius <= 4096.) || !(b->brush_opacity >= 0. && b->brush_opacity <= 1.)) FAIL("brush radius / opacity out of range");
  if (b->winding != COH_BRUSH_GAUSSIAN && b->winding != COH_BRUSH_DUMMY) FAIL("bad brush kind");
  *r = b->winding == COH_BRUSH_DUMMY ? (int)b->brush_radius : (int)ceil(b->brush_radius);   // ((2 r + 1) - 1) / 2 of sizeof_brush (brush.ml:25-28)
  return 0;
}
int coh_brush_shape(coh_ctx* ctx, const coh_object* brush, const int32_t* points, int32_t n_points, coh_shape_t* shape) {
  CK(cudaSetDevice(ctx->device));
  *shape = 0;
  int r;
  if (brush_radius_of(ctx, brush, &r)) return 1;
  if (n_points <= 0) return 0;   // NullShape (brush.ml:155)
  int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
  for (int k = 0; k < n_points; k++) { x0 = std::min(x0, points[2 * k]); x1 = std::max(x1, points[2 * k]); y0 = std::min(y0, points[2 * k + 1]); y1 = std::max(y1, points[2 * k + 1]); }
  x0 -= r; y0 -= r; x1 += r; y1 += r;
  const int wx0 = floordiv(x0, 32) * 32, nw = (x1 - wx0) / 32 + 1, n_rows = y1 - y0 + 1;
  int2* d_pts = nullptr; uint32_t* bits = nullptr;
  CK(DMALLOC(&d_pts, sizeof(int2) * (size_t)n_points)); CK(DMALLOC(&bits, 4 * (size_t)nw * n_rows));
  CK(cudaMemcpyAsync(d_pts, points, sizeof(int2) * (size_t)n_points, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(bits, 0, 4 * (size_t)nw * n_rows, ctx->stream));
  const int side = 2 * r + 1;
  k_stamp_boxes_to_bits<<<cdiv(n_points * side, 256), 256, 0, ctx->stream>>>(d_pts, n_points, r, y0, n_rows, wx0, nw, bits); LAUNCHED();
  int rc = shape_from_bits(ctx, bits, y0, n_rows, wx0, nw, shape);
  CK(cudaStreamSynchronize(ctx->stream));   // (`points` is the caller's)
  DFREE(d_pts); DFREE(bits);
  return rc;
}
// device copy of a brush's stamp (alpha bytes of Brush.drawbrush brush white, brush.ml:75-92)
static int upload_stamp(coh_ctx* ctx, const coh_object* brush, int r, uint8_t** d_stamp) {
  std::vector<uint8_t> st; int rr = r;
  if (brush->winding == COH_BRUSH_DUMMY) st.assign((size_t)(2 * r + 1) * (2 * r + 1), (uint8_t)255);
  else brush_stamp(brush->brush_radius, brush->brush_opacity, st, rr, nullptr);
  CK(DMALLOC(d_stamp, st.size()));
  CK(cudaMemcpyAsync(*d_stamp, st.data(), st.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int coh_brush_sprite(coh_ctx* ctx, const coh_object* brush, const int32_t* points, int32_t n_points, coh_shape_t shp,
                     uint32_t* out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  int r;
  if (brush_radius_of(ctx, brush, &r)) return 1;
  if (brush->fill_kind < COH_FILL_PLAIN || brush->fill_kind > COH_FILL_RADIAL) FAIL("coh_brush_sprite: bad fill kind");
  if (!shp) return 0;   // NullShape -> NullSprite (brush.ml:184)
  DevShape* s = (DevShape*)shp;
  if (s->card > cap) FAIL("coh_brush_sprite: buffer too small");
  int2* d_pts = nullptr; uint8_t* d_stamp = nullptr; int* d_off = nullptr; uint32_t* d_out = nullptr;
  CK(DMALLOC(&d_pts, sizeof(int2) * (size_t)std::max(n_points, 1)));
  if (n_points > 0) CK(cudaMemcpyAsync(d_pts, points, sizeof(int2) * (size_t)n_points, cudaMemcpyHostToDevice, ctx->stream));
  if (upload_stamp(ctx, brush, r, &d_stamp)) return 1;
  if (shape_pixel_offsets(ctx, s, &d_off)) return 1;
  CK(DMALLOC(&d_out, 4 * (size_t)std::max<long long>(s->card, 1)));
  FillRec f = fillrec_of(brush);
  if (brush->winding == COH_BRUSH_DUMMY) { f.kind = 0; f.c0 = 0xFFFFFFFFu; }   // white, whatever the fill (brush.ml:178-181)
  k_brush_sprite<<<cdiv(s->n_rows, 4), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, d_off, s->y0, s->n_rows, d_pts, n_points, r, d_stamp, f, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(out, d_out, 4 * (size_t)s->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(d_pts); DFREE(d_stamp); DFREE(d_off); DFREE(d_out);
  *n_out = s->card;
  return 0;
}
// Brush.smear spr brushstroke (brush.ml:286-331): the sprite fleshed out to the stroke's shape, smeared along the smear
// points (coh_host_smear_points) on a canvas of its box plus one pixel; result = a sprite on shape(spr) ∪ shape(stroke).
int coh_brush_smear(coh_ctx* ctx, coh_shape_t shape, const uint32_t* rgba_in, const coh_object* brush, const int32_t* points, int32_t n_points,
                    const int32_t* smear_points, int32_t n_smear, coh_shape_t* out_shape, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *out_shape = 0; *n_out = 0;
  int r;
  if (brush_radius_of(ctx, brush, &r)) return 1;
  if (brush->winding != COH_BRUSH_GAUSSIAN) FAIL("Brush.drawbrush : One cannot draw a dummy brush");   // brush.ml:98-99
  if (r > 64) FAIL("coh_brush_smear: brush radius above 64");
  coh_shape_t bs = 0, U = 0;
  if (coh_brush_shape(ctx, brush, points, n_points, &bs)) return 1;
  if (coh_shape_union(ctx, shape, bs, &U)) return 1;
  DevShape* s = (DevShape*)shape; DevShape* us = (DevShape*)U; DevShape* bsh = (DevShape*)bs;
  if (!us) { coh_shape_free(ctx, bs); return 0; }   // NullSprite
  if (us->card > cap) { coh_shape_free(ctx, bs); coh_shape_free(ctx, U); FAIL("coh_brush_smear: buffer too small"); }
  // canvas coordinates: the box of the fleshed-out sprite with a border of one pixel (Sprite.flatten_sprite 1)
  const int ax0 = us->bx0 - 1, ay0 = us->by0 - 1, cw = us->bx1 - us->bx0 + 3, ch = us->by1 - us->by0 + 3;
  const size_t npx = (size_t)cw * ch;
  uint32_t *A = nullptr, *Cv = nullptr, *Y = nullptr, *d_in = nullptr, *d_out = nullptr; int *d_off = nullptr, *d_uoff = nullptr, *d_bb = nullptr;
  int2* d_sm = nullptr; uint8_t* d_stamp = nullptr;
  CK(DMALLOC(&A, 4 * npx)); CK(DMALLOC(&Cv, 4 * npx)); CK(DMALLOC(&Y, 4 * npx)); CK(DMALLOC(&d_bb, 4 * sizeof(int)));
  CK(cudaMemsetAsync(A, 0, 4 * npx, ctx->stream));
  if (s) {
    if (shape_pixel_offsets(ctx, s, &d_off)) return 1;
    CK(DMALLOC(&d_in, 4 * (size_t)std::max<long long>(s->card, 1)));
    CK(cudaMemcpyAsync(d_in, rgba_in, 4 * (size_t)s->card, cudaMemcpyHostToDevice, ctx->stream));
    k_scatter_spans<uint32_t><<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, d_off, s->n_rows, s->y0 - ay0, ax0, cw, d_in, A); LAUNCHED();
  }
  // the box of the fleshed-out sprite is the box of the union: k_smear takes it as (sprite box, stroke box)
  const int bb[4] = {us->bx0 - ax0, us->by0 - ay0, us->bx1 - ax0, us->by1 - ay0};
  CK(cudaMemcpyAsync(d_bb, bb, sizeof bb, cudaMemcpyHostToDevice, ctx->stream));
  CK(DMALLOC(&d_sm, sizeof(int2) * (size_t)std::max(n_smear, 1)));
  if (n_smear > 0) CK(cudaMemcpyAsync(d_sm, smear_points, sizeof(int2) * (size_t)n_smear, cudaMemcpyHostToDevice, ctx->stream));
  if (upload_stamp(ctx, brush, r, &d_stamp)) return 1;
  k_smear<<<1, SMEAR_THREADS, 0, ctx->stream>>>(A, Y, cw, ch, Cv, 0, 0, cw, ch, d_bb, bb[0], bb[1], bb[2], bb[3], d_sm, n_smear, -ax0, -ay0, d_stamp, r, 0, 0, cw - 1, ch - 1); LAUNCHED();
  // Sprite.pickup on the fleshed-out shape
  if (shape_pixel_offsets(ctx, us, &d_uoff)) return 1;
  CK(DMALLOC(&d_out, 4 * (size_t)std::max<long long>(us->card, 1)));
  k_gather_spans<uint32_t><<<cdiv(us->n_rows, 128), 128, 0, ctx->stream>>>(us->row_ptr, us->spans, d_uoff, us->n_rows, ax0, cw, Y + (size_t)(us->y0 - ay0) * cw, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)us->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(A); DFREE(Cv); DFREE(Y); DFREE(d_in); DFREE(d_out); DFREE(d_off); DFREE(d_uoff); DFREE(d_bb); DFREE(d_sm); DFREE(d_stamp);
  (void)bsh;
  coh_shape_free(ctx, bs);
  *out_shape = U; *n_out = us->card;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Sprite operations on whole sprites (sprite.mli:96-125).  A sprite crosses the boundary as its shape (a device span
// set) and one RGBA8 word per pixel in canonical span order.  Sprite.translate_sprite moves only the shape
// (coh_shape_translate): the pixel array is unchanged.
// ---------------------------------------------------------------------------------------
static FillRec fillrec_of(const coh_object* fill) {
  FillRec f; f.kind = fill->fill_kind; f.c0 = fill->colour0; f.c1 = fill->colour1; f.flags = fill->fill_flags;
  for (int i = 0; i < 6; i++) f.p[i] = fill->fparam[i];
  return f;
}
int coh_shape_intersects(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, int32_t* yes) {   // sprite.ml:1661-1662
  coh_shape_t t = 0;
  *yes = 0;
  if (coh_shape_intersection(ctx, a, b, &t)) return 1;
  *yes = t != 0;
  return coh_shape_free(ctx, t);
}
// Sprite.portion spr shp (sprite.ml:642-721): the pixels of the sprite on `sub`, which must lie inside its shape
int coh_sprite_portion(coh_ctx* ctx, coh_shape_t shape, const uint32_t* rgba, coh_shape_t sub, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  DevShape* a = (DevShape*)shape; DevShape* b = (DevShape*)sub;
  if (!b) return 0;
  if (!a) FAIL("portion: malformed input (sprite null, shape not)");
  if (b->card > cap) FAIL("coh_sprite_portion: buffer too small");
  coh_shape_t outside = 0;
  if (coh_shape_difference(ctx, sub, shape, &outside)) return 1;
  if (outside) { coh_shape_free(ctx, outside); FAIL("portion_spanline: bad input"); }   // shp is not a subset of the sprite's shape
  const int wx0 = floordiv(a->bx0, 32) * 32, w = a->bx1 - wx0 + 1, h = a->n_rows;
  uint32_t *canvas = nullptr, *d_in = nullptr, *d_out = nullptr; int *offa = nullptr, *offb = nullptr;
  if (shape_pixel_offsets(ctx, a, &offa) || shape_pixel_offsets(ctx, b, &offb)) return 1;
  CK(DMALLOC(&canvas, 4 * (size_t)w * h)); CK(DMALLOC(&d_in, 4 * (size_t)a->card)); CK(DMALLOC(&d_out, 4 * (size_t)b->card));
  CK(cudaMemcpyAsync(d_in, rgba, 4 * (size_t)a->card, cudaMemcpyHostToDevice, ctx->stream));
  k_scatter_spans<uint32_t><<<cdiv(a->n_rows, 128), 128, 0, ctx->stream>>>(a->row_ptr, a->spans, offa, a->n_rows, 0, wx0, w, d_in, canvas); LAUNCHED();
  k_gather_spans<uint32_t><<<cdiv(b->n_rows, 128), 128, 0, ctx->stream>>>(b->row_ptr, b->spans, offb, b->n_rows, wx0, w, canvas + (size_t)(b->y0 - a->y0) * w, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)b->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(canvas); DFREE(d_in); DFREE(d_out); DFREE(offa); DFREE(offb);
  *n_out = b->card;
  return 0;
}
// Sprite.fillshape shp fill (sprite.ml:158-175)
int coh_sprite_fillshape(coh_ctx* ctx, coh_shape_t shape, const coh_object* fill, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  DevShape* s = (DevShape*)shape;
  if (!s) return 0;
  if (s->card > cap) FAIL("coh_sprite_fillshape: buffer too small");
  if (fill->fill_kind < COH_FILL_PLAIN || fill->fill_kind > COH_FILL_RADIAL) FAIL("coh_sprite_fillshape: bad fill kind");
  int* off = nullptr; uint32_t* d_out = nullptr;
  if (shape_pixel_offsets(ctx, s, &off)) return 1;
  CK(DMALLOC(&d_out, 4 * (size_t)s->card));
  k_sprite_fillshape<<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, off, s->y0, s->n_rows, fillrec_of(fill), d_out); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)s->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(off); DFREE(d_out);
  *n_out = s->card;
  return 0;
}
// Sprite.sprite_map f spr with f = Colour.monochrome | dissolve ~delta:arg | red / green / blue_channel
int coh_sprite_map(coh_ctx* ctx, int32_t op, int32_t arg, const uint32_t* rgba_in, int64_t n, uint32_t* rgba_out) {
  CK(cudaSetDevice(ctx->device));
  if (op < COH_MAP_MONOCHROME || op > COH_MAP_BLUE_CHANNEL) FAIL("coh_sprite_map: unknown colour function");
  if (op == COH_MAP_DISSOLVE && (arg < 0 || arg > 255)) FAIL("Colour.dissolve: delta out of range");   // colour.ml:292 assert
  if (n <= 0) return 0;
  uint32_t *d_in = nullptr, *d_out = nullptr;
  CK(DMALLOC(&d_in, 4 * (size_t)n)); CK(DMALLOC(&d_out, 4 * (size_t)n));
  CK(cudaMemcpyAsync(d_in, rgba_in, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  k_sprite_map<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(op, arg, d_in, d_out, (size_t)n); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(d_in); DFREE(d_out);
  return 0;
}
// Sprite.map_coords (fun x y c -> Colour.dissolve (fill x y) ~delta:(alpha c)) spr: a fill applied to an alpha matte,
// the closing step of Render.sprite_of_cpg (render.ml:976-981)
int coh_sprite_map_coords_fill(coh_ctx* ctx, coh_shape_t shape, const coh_object* fill, const uint32_t* rgba_in, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  DevShape* s = (DevShape*)shape;
  if (!s) return 0;
  if (s->card > cap) FAIL("coh_sprite_map_coords_fill: buffer too small");
  if (fill->fill_kind < COH_FILL_PLAIN || fill->fill_kind > COH_FILL_RADIAL) FAIL("coh_sprite_map_coords_fill: bad fill kind");
  int* off = nullptr; uint32_t *d_in = nullptr, *d_out = nullptr;
  if (shape_pixel_offsets(ctx, s, &off)) return 1;
  CK(DMALLOC(&d_in, 4 * (size_t)s->card)); CK(DMALLOC(&d_out, 4 * (size_t)s->card));
  CK(cudaMemcpyAsync(d_in, rgba_in, 4 * (size_t)s->card, cudaMemcpyHostToDevice, ctx->stream));
  k_sprite_fill_alpha<<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, off, s->y0, s->n_rows, fillrec_of(fill), d_in, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)s->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(off); DFREE(d_in); DFREE(d_out);
  *n_out = s->card;
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
// Convolve.convolve_sprite kernel sprite (convolve.ml:239-258): the sprite is (shape, one RGBA8 per pixel in
// span order); the result lives on bloat r r (shape) and is returned the same way.
int coh_convolve_sprite(coh_ctx* ctx, int32_t kernel_kind, int32_t r, coh_shape_t shape, const uint32_t* rgba_in,
                        coh_shape_t* out_shape, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *out_shape = 0; *n_out = 0;
  if ((kernel_kind != COH_CONV_UNIT && kernel_kind != COH_CONV_GAUSSIAN) || r <= 0) FAIL("Convolve.mkunit / Convolve.mkxy: Invalid_argument");
  if (r > 64) FAIL("coh_convolve_sprite: kernel radius above 64 (the 32-bit tap sums of k_conv_pass are sized for it)");
  if (!shape) return 0;  // NullSprite -> NullSprite
  DevShape* s = (DevShape*)shape;
  coh_shape_t R = 0;
  if (coh_shape_bloat(ctx, shape, r, r, &R)) return 1;
  DevShape* rs = (DevShape*)R;
  if (rs->card > cap) { coh_shape_free(ctx, R); FAIL("coh_convolve_sprite: buffer too small"); }
  // canvas = bounding box grown by 2r (Sprite.flatten_sprite border, convolve.ml:247)
  const int x0 = s->bx0 - 2 * r, y0 = s->by0 - 2 * r, w = s->bx1 - s->bx0 + 1 + 4 * r, h = s->by1 - s->by0 + 1 + 4 * r;
  const size_t npx = (size_t)w * h;
  uint32_t *A = nullptr, *X = nullptr, *d_in = nullptr, *d_out = nullptr; int* d_off = nullptr; int* d_taps = nullptr; int total = 0;
  if (shape_pixel_offsets(ctx, s, &d_off)) return 1;
  CK(DMALLOC(&A, 4 * npx)); CK(DMALLOC(&X, 4 * npx));
  CK(cudaMemsetAsync(A, 0, 4 * npx, ctx->stream));
  CK(DMALLOC(&d_in, 4 * (size_t)std::max<long long>(s->card, 1)));
  CK(cudaMemcpyAsync(d_in, rgba_in, 4 * (size_t)s->card, cudaMemcpyHostToDevice, ctx->stream));
  k_scatter_spans<uint32_t><<<cdiv(s->n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, d_off, s->n_rows, s->y0 - y0, x0, w, d_in, A); LAUNCHED();
  if (conv_taps(ctx, kernel_kind, r, &d_taps, &total)) return 1;
  dim3 gp(cdiv(w, 128), h);
  k_conv_pass<<<gp, 128, 0, ctx->stream>>>(A, X, w, h, r, kernel_kind, d_taps, total, 0); LAUNCHED();
  k_conv_pass<<<gp, 128, 0, ctx->stream>>>(X, A, w, h, r, kernel_kind, d_taps, total, 1); LAUNCHED();
  // pick the result up on R (Sprite.pickup)
  int* d_roff = nullptr;
  if (shape_pixel_offsets(ctx, rs, &d_roff)) return 1;
  CK(DMALLOC(&d_out, 4 * (size_t)std::max<long long>(rs->card, 1)));
  // k_gather_spans indexes dense rows from the shape's first row: pass the canvas rows starting at R's first row
  k_gather_spans<uint32_t><<<cdiv(rs->n_rows, 128), 128, 0, ctx->stream>>>(rs->row_ptr, rs->spans, d_roff, rs->n_rows, x0, w, A + (size_t)(rs->y0 - y0) * w, d_out); LAUNCHED();
  CK(cudaMemcpyAsync(rgba_out, d_out, 4 * (size_t)rs->card, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(A); DFREE(X); DFREE(d_in); DFREE(d_out); DFREE(d_off); DFREE(d_roff); DFREE(d_taps);
  *out_shape = R; *n_out = rs->card;
  return 0;
}
