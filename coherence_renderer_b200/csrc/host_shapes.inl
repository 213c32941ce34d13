// host_shapes.inl — part of coherence_b200.cu (one translation unit; included in order): span sets (sprite.mli shapes): device CSR <-> bit-frames, set algebra, bloat / erode.

// ---------------------------------------------------------------------------------------
// Shapes
// ---------------------------------------------------------------------------------------
static void free_shape(coh_ctx* ctx, DevShape* s) {
  if (!s) return;
  DFREE(s->row_ptr); DFREE(s->spans);
  delete s;
}
int coh_shape_free(coh_ctx* ctx, coh_shape_t h) {
  CK(cudaSetDevice(ctx->device));
  free_shape(ctx, (DevShape*)h);
  return 0;
}

// Bit-frame [n_rows][nw] (device) -> span set.  Consumes nothing; returns 0 handle for the empty set.
static int shape_from_bits(coh_ctx* ctx, const uint32_t* bits, int y0, int n_rows, int wx0, int nw, coh_shape_t* out) {
  *out = 0;
  if (n_rows <= 0 || nw <= 0) return 0;
  int* counts = nullptr; int* ptr = nullptr; ShapeMeta* d_meta = nullptr;
  CK(DMALLOC(&counts, sizeof(int) * n_rows));
  CK(DMALLOC(&ptr, sizeof(int) * (n_rows + 1)));
  CK(DMALLOC(&d_meta, sizeof(ShapeMeta)));
  const ShapeMeta init{0ull, 0, 0x7FFFFFFF, -1, 0x7FFFFFFF, -1, 0};
  CK(cudaMemcpyAsync(d_meta, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  k_count_runs<<<cdiv(n_rows, 128), 128, 0, ctx->stream>>>(bits, n_rows, nw, counts, d_meta); LAUNCHED();
  if (exclusive_scan(ctx, counts, ptr, n_rows, nullptr)) return 1;
  // one small read-back: the host keeps cardinality, span count and tight bounds of every span set (they size the
  // bit-frames of later operations and answer coh_shape_bounds / coh_shape_card); the spans stay on the device
  ShapeMeta m;
  CK(cudaMemcpyAsync(&m, d_meta, sizeof m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  DFREE(counts); DFREE(d_meta);
  if (m.n_spans == 0) { DFREE(ptr); return 0; }
  const int first = m.first_row, last = m.last_row;   // empty rows at both ends are trimmed: y0 / n_rows are tight
  DevShape* s = new DevShape();
  s->n_spans = m.n_spans; s->card = (long long)m.card;
  CK(DMALLOC(&s->spans, sizeof(int2) * m.n_spans));
  k_fill_runs<<<cdiv(n_rows, 128), 128, 0, ctx->stream>>>(bits, n_rows, nw, wx0, ptr, s->spans); LAUNCHED();
  s->y0 = y0 + first; s->n_rows = last - first + 1;
  CK(DMALLOC(&s->row_ptr, sizeof(int) * (s->n_rows + 1)));
  CK(cudaMemcpyAsync(s->row_ptr, ptr + first, sizeof(int) * (s->n_rows + 1), cudaMemcpyDeviceToDevice, ctx->stream));
  DFREE(ptr);
  s->by0 = s->y0; s->by1 = s->y0 + s->n_rows - 1; s->bx0 = wx0 + m.bit_lo; s->bx1 = wx0 + m.bit_hi;
  *out = (coh_shape_t)s;
  return 0;
}
// span set -> freshly allocated bit-frame covering rows [y0, y0+n_rows) and words from pixel wx0
static int bits_from_shape(coh_ctx* ctx, const DevShape* s, int y0, int n_rows, int wx0, int nw, uint32_t** out) {
  uint32_t* bits = nullptr;
  CK(DMALLOC(&bits, sizeof(uint32_t) * (size_t)n_rows * nw));
  CK(cudaMemsetAsync(bits, 0, sizeof(uint32_t) * (size_t)n_rows * nw, ctx->stream));
  if (s && s->n_spans > 0) {
    k_spans_to_bits<<<cdiv(n_rows, 128), 128, 0, ctx->stream>>>(s->row_ptr, s->spans, s->y0, s->n_rows, y0, n_rows, wx0, nw, bits);
    LAUNCHED();
  }
  *out = bits;
  return 0;
}

int coh_shape_box(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (w == 0 && h == 0) return 0;                       // sprite.ml:463
  if (w < 0 || h < 0) FAIL("Sprite.box: negative argument.");  // sprite.ml:464
  if (w == 0 || h == 0) return 0;
  // Sprite.box: h rows of one span (x, w), written by a kernel — nothing is built on the host
  DevShape* s = new DevShape();
  s->y0 = y; s->n_rows = h; s->n_spans = h; s->card = (long long)w * h;
  s->bx0 = x; s->bx1 = x + w - 1; s->by0 = y; s->by1 = y + h - 1;
  CK(DMALLOC(&s->row_ptr, sizeof(int) * (h + 1)));
  CK(DMALLOC(&s->spans, sizeof(int2) * h));
  k_box_spans<<<cdiv(h + 1, 256), 256, 0, ctx->stream>>>(s->row_ptr, s->spans, h, x, w); LAUNCHED();
  *out = (coh_shape_t)s;
  return 0;
}
int coh_shape_import(coh_ctx* ctx, const int32_t* flat, int64_t n, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (n == 0) return 0;
  // validate canonical form (sprite.ml:201-239) while building the CSR
  std::vector<int> ys; std::vector<int> cnt; std::vector<int2> spans;
  int64_t i = 0; long long card = 0;
  int bx0 = INT32_MAX, bx1 = INT32_MIN;
  while (i < n) {
    if (i + 2 > n) FAIL("shape import: truncated row header");
    int y = flat[i], k = flat[i + 1]; i += 2;
    if (k <= 0) FAIL("shape import: malformed shape (empty spanline)");
    if (!ys.empty() && y <= ys.back()) FAIL("shape import: malformed shape (rows not increasing)");
    if (i + 2 * (int64_t)k > n) FAIL("shape import: truncated spans");
    for (int q = 0; q < k; q++, i += 2) {
      int x = flat[i], l = flat[i + 1];
      if (l <= 0) FAIL("shape import: malformed shape (span length)");
      if (q && x <= spans.back().x + spans.back().y) FAIL("shape import: malformed shape (spans overlap or abut)");
      spans.push_back(make_int2(x, l)); card += l;
      bx0 = std::min(bx0, x); bx1 = std::max(bx1, x + l - 1);
    }
    ys.push_back(y); cnt.push_back(k);
  }
  DevShape* s = new DevShape();
  s->y0 = ys.front(); s->n_rows = ys.back() - ys.front() + 1;
  std::vector<int> ptr(s->n_rows + 1, 0);
  for (size_t r = 0; r < ys.size(); r++) ptr[ys[r] - s->y0 + 1] = cnt[r];
  for (int r = 0; r < s->n_rows; r++) ptr[r + 1] += ptr[r];
  s->n_spans = (int)spans.size(); s->card = card;
  s->bx0 = bx0; s->bx1 = bx1; s->by0 = ys.front(); s->by1 = ys.back();
  CK(DMALLOC(&s->row_ptr, sizeof(int) * ptr.size()));
  CK(DMALLOC(&s->spans, sizeof(int2) * spans.size()));
  CK(cudaMemcpyAsync(s->row_ptr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(s->spans, spans.data(), sizeof(int2) * spans.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *out = (coh_shape_t)s;
  return 0;
}
static int download_shape(coh_ctx* ctx, const DevShape* s, std::vector<int>& ptr, std::vector<int2>& spans) {
  ptr.resize(s->n_rows + 1); spans.resize(s->n_spans);
  CK(cudaMemcpyAsync(ptr.data(), s->row_ptr, sizeof(int) * ptr.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(spans.data(), s->spans, sizeof(int2) * spans.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int coh_shape_export_size(coh_ctx* ctx, coh_shape_t h, int64_t* n) {
  CK(cudaSetDevice(ctx->device));
  *n = 0;
  if (!h) return 0;
  DevShape* s = (DevShape*)h;
  std::vector<int> ptr; std::vector<int2> spans;
  if (download_shape(ctx, s, ptr, spans)) return 1;
  int64_t rows = 0;
  for (int r = 0; r < s->n_rows; r++) rows += ptr[r + 1] > ptr[r];
  *n = 2 * rows + 2 * (int64_t)s->n_spans;
  return 0;
}
int coh_shape_export(coh_ctx* ctx, coh_shape_t h, int32_t* flat, int64_t cap, int64_t* n_out) {
  CK(cudaSetDevice(ctx->device));
  *n_out = 0;
  if (!h) return 0;
  DevShape* s = (DevShape*)h;
  std::vector<int> ptr; std::vector<int2> spans;
  if (download_shape(ctx, s, ptr, spans)) return 1;
  int64_t k = 0;
  for (int r = 0; r < s->n_rows; r++) {
    int c = ptr[r + 1] - ptr[r];
    if (!c) continue;
    if (k + 2 + 2 * c > cap) FAIL("coh_shape_export: buffer too small");
    flat[k++] = s->y0 + r; flat[k++] = c;
    for (int q = ptr[r]; q < ptr[r + 1]; q++) { flat[k++] = spans[q].x; flat[k++] = spans[q].y; }
  }
  *n_out = k;
  return 0;
}
int coh_shape_bounds(coh_ctx* ctx, coh_shape_t h, int32_t box[4], int32_t* is_null) {
  (void)ctx;
  DevShape* s = (DevShape*)h;
  *is_null = !s;
  if (s) { box[0] = s->bx0; box[1] = s->by0; box[2] = s->bx1; box[3] = s->by1; }
  return 0;
}
int coh_shape_card(coh_ctx* ctx, coh_shape_t h, int64_t* n) {
  (void)ctx;
  *n = h ? ((DevShape*)h)->card : 0;
  return 0;
}

// Binary set algebra through bit-frames over the union bounding box (K3).
static int shape_binop(coh_ctx* ctx, coh_shape_t ha, coh_shape_t hb, int op, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  DevShape* a = (DevShape*)ha; DevShape* b = (DevShape*)hb;
  if (!a && !b) return 0;
  if (!a && op != 0) return 0;       // {} - b = {} ; {} & b = {}
  if (!b && op == 2) return 0;
  int x0 = INT32_MAX, x1 = INT32_MIN, y0 = INT32_MAX, y1 = INT32_MIN;
  for (DevShape* s : {a, b}) if (s) { x0 = std::min(x0, s->bx0); x1 = std::max(x1, s->bx1); y0 = std::min(y0, s->by0); y1 = std::max(y1, s->by1); }
  int wx0 = floordiv(x0, 32) * 32, nw = (x1 - wx0) / 32 + 1, n_rows = y1 - y0 + 1;
  uint32_t *ba = nullptr, *bb = nullptr;
  if (bits_from_shape(ctx, a, y0, n_rows, wx0, nw, &ba)) return 1;
  if (bits_from_shape(ctx, b, y0, n_rows, wx0, nw, &bb)) return 1;
  size_t n = (size_t)n_rows * nw;
  k_bitop<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ba, bb, ba, n, op); LAUNCHED();
  int rc = shape_from_bits(ctx, ba, y0, n_rows, wx0, nw, out);
  DFREE(ba); DFREE(bb);
  return rc;
}
int coh_shape_union(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out) { return shape_binop(ctx, a, b, 0, out); }
int coh_shape_difference(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out) { return shape_binop(ctx, a, b, 1, out); }
int coh_shape_intersection(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out) { return shape_binop(ctx, a, b, 2, out); }

int coh_shape_translate(coh_ctx* ctx, coh_shape_t h, int32_t dx, int32_t dy, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (!h) return 0;
  DevShape* s = (DevShape*)h;
  DevShape* t = new DevShape(*s);
  t->y0 += dy; t->bx0 += dx; t->bx1 += dx; t->by0 += dy; t->by1 += dy;
  CK(DMALLOC(&t->row_ptr, sizeof(int) * (s->n_rows + 1)));
  CK(DMALLOC(&t->spans, sizeof(int2) * s->n_spans));
  CK(cudaMemcpyAsync(t->row_ptr, s->row_ptr, sizeof(int) * (s->n_rows + 1), cudaMemcpyDeviceToDevice, ctx->stream));
  k_translate_spans<<<cdiv(s->n_spans, 256), 256, 0, ctx->stream>>>(s->spans, t->spans, s->n_spans, dx); LAUNCHED();
  *out = (coh_shape_t)t;
  return 0;
}
static int bloat_impl(coh_ctx* ctx, const DevShape* s, int x0, int y0, int x1, int y1, int m, int n, bool complement_in_box,
                      coh_shape_t* out) {
  // frame = box [x0..x1] x [y0..y1] grown by (m, n) on every side
  int fx0 = x0 - m, fy0 = y0 - n, fx1 = x1 + m, fy1 = y1 + n;
  int wx0 = floordiv(fx0, 32) * 32, nw = (fx1 - wx0) / 32 + 1, n_rows = fy1 - fy0 + 1;
  uint32_t *in = nullptr, *tmp = nullptr;
  if (bits_from_shape(ctx, s, fy0, n_rows, wx0, nw, &in)) return 1;
  CK(DMALLOC(&tmp, sizeof(uint32_t) * (size_t)n_rows * nw));
  size_t nwords = (size_t)n_rows * nw;
  if (complement_in_box) {
    // erode (sprite.ml:1867-1877): inverse = enclosing - shp, bloated, then shp - bloated
    uint32_t* box = nullptr;
    CK(DMALLOC(&box, sizeof(uint32_t) * nwords));
    dim3 g(cdiv(nw, 128), n_rows);
    k_fill_box_bits<<<g, 128, 0, ctx->stream>>>(box, n_rows, nw, wx0, fy0, fx0, fy0, fx1, fy1); LAUNCHED();
    k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(box, in, box, nwords, 1); LAUNCHED();  // inverse
    k_dilate<<<g, 128, 0, ctx->stream>>>(box, tmp, n_rows, nw, m, n); LAUNCHED();
    k_bitop<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(in, tmp, tmp, nwords, 1); LAUNCHED();    // shp - bloated
    DFREE(box);
  } else {
    dim3 g(cdiv(nw, 128), n_rows);
    k_dilate<<<g, 128, 0, ctx->stream>>>(in, tmp, n_rows, nw, m, n); LAUNCHED();
  }
  int rc = shape_from_bits(ctx, tmp, fy0, n_rows, wx0, nw, out);
  DFREE(in); DFREE(tmp);
  return rc;
}
int coh_shape_bloat(coh_ctx* ctx, coh_shape_t h, int32_t m, int32_t n, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (!h) return 0;
  if (m < 0 || n < 0) FAIL("Sprite.bloat: negative radius");
  DevShape* s = (DevShape*)h;
  return bloat_impl(ctx, s, s->bx0, s->by0, s->bx1, s->by1, m, n, false, out);
}
int coh_shape_erode(coh_ctx* ctx, coh_shape_t h, int32_t m, int32_t n, coh_shape_t* out) {
  CK(cudaSetDevice(ctx->device));
  *out = 0;
  if (!h) return 0;
  if (m < 0 || n < 0) FAIL("Sprite.erode: negative radius");
  DevShape* s = (DevShape*)h;
  return bloat_impl(ctx, s, s->bx0, s->by0, s->bx1, s->by1, m, n, true, out);
}
