// kernels_shapes.cuh — part of kernels.cuh (included inside namespace coh, in order): span-set algebra on bit-frames, filter and convolve passes, exports.

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
// The same span set at two positions into one bit-frame (shp_o ||| shp_n of a drag step, render.ml:1396-1400): one launch
// instead of two, one WARP per destination row (k_spans_to_bits' thread per row walks the row's spans one after the
// other, each a read-modify-write in global memory: 11 us for the lion group).  Lanes take the spans of either
// position side by side and OR their words atomically (the bit-frame is zeroed by the caller).
__global__ void k_spans_to_bits2(const int* __restrict__ row_ptr, const int2* __restrict__ spans, int src_y0a, int wx0a, int src_y0b, int wx0b,
                                 int src_rows, int n_rows, int nw, uint32_t* __restrict__ bits) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;  // destination row
  if (r >= n_rows) return;
  uint32_t* row = bits + (size_t)r * nw;
  const int nb = nw * 32;
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int sr = r - (pass ? src_y0b : src_y0a), wx0 = pass ? wx0b : wx0a;
    if (sr < 0 || sr >= src_rows) continue;
    const int k1 = row_ptr[sr + 1];
    for (int k = row_ptr[sr] + lane; k < k1; k += 32) {
      const int2 s = spans[k];
      int a = s.x - wx0, b = s.x + s.y - 1 - wx0;   // (or_interval's clipping, raster_core.cuh)
      if (b < 0 || a >= nb) continue;
      a = max(a, 0); b = min(b, nb - 1);
      if (a > b) continue;
      const int wa = a >> 5, wb = b >> 5;
      const uint32_t ma = 0xFFFFFFFFu << (a & 31), mb = 0xFFFFFFFFu >> (31 - (b & 31));
      if (wa == wb) { atomicOr(&row[wa], ma & mb); continue; }
      atomicOr(&row[wa], ma);
      for (int w = wa + 1; w < wb; w++) atomicOr(&row[w], 0xFFFFFFFFu);
      atomicOr(&row[wb], mb);
    }
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
// The same dilation for m <= 32: a row's horizontal part comes from three words by shift doubling, so a word costs
// O((2n+1) log m) operations instead of O((2n+1)(2m+1)) (the blur filter's reading shape, filters.ml:247: m = n = 2r+1).
__device__ __forceinline__ uint32_t hdilate_word(uint32_t a, uint32_t b, uint32_t c, int m) {
  // bits of b's word: pixel i set when any of pixels i-m .. i+m is (a = the word to the left, c = to the right)
  unsigned long long lo = ((unsigned long long)b << 32) | a;   // smear towards higher pixels, read the high half
  unsigned long long hi = ((unsigned long long)c << 32) | b;   // smear towards lower pixels, read the low half
  for (int cover = 0; cover < m;) {
    const int step = min(cover + 1, m - cover);
    lo |= lo << step; hi |= hi >> step;
    cover += step;
  }
  return (uint32_t)(lo >> 32) | (uint32_t)hi;
}
__global__ void k_dilate32(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n_rows, int nw, int m, int n) {
  int w = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (w >= nw || r >= n_rows) return;
  uint32_t acc = 0u;
  for (int rr = max(0, r - n); rr <= min(n_rows - 1, r + n); rr++) {
    const uint32_t* row = in + (size_t)rr * nw;
    acc |= hdilate_word(w > 0 ? row[w - 1] : 0u, row[w], w + 1 < nw ? row[w + 1] : 0u, m);
  }
  out[(size_t)r * nw + w] = acc;
}
// Sprite.box x y w h (sprite.ml:462-465) as a device span set: every row one span
__global__ void k_box_spans(int* __restrict__ row_ptr, int2* __restrict__ spans, int h, int x, int w) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > h) return;
  row_ptr[r] = r;
  if (r < h) spans[r] = make_int2(x, w);
}
// What the host needs to know of a span set, reduced on the device (one small read-back instead of the spans):
struct ShapeMeta { unsigned long long card; int n_spans, first_row, last_row, bit_lo, bit_hi, pad; };
// Run extraction: count the maximal runs of every row (thread per row), then fill.  `meta` (optional) receives the
// cardinality, the number of runs, the first / last non-empty row and the lowest / highest set bit of any row.
__global__ void k_count_runs(const uint32_t* __restrict__ bits, int n_rows, int nw, int* __restrict__ counts,
                             ShapeMeta* __restrict__ meta) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint32_t* row = bits + (size_t)r * nw;
  int n = 0; uint32_t carry = 0u; unsigned long long px = 0;
  int lo = 0x7FFFFFFF, hi = -1;
  for (int w = 0; w < nw; w++) {
    uint32_t v = row[w];
    n += __popc(v & ~((v << 1) | carry));
    px += __popc(v);
    carry = v >> 31;
    if (v) { if (hi < 0) lo = 32 * w + __ffs((int)v) - 1; hi = 32 * w + 31 - __clz((int)v); }
  }
  counts[r] = n;
  if (meta && px) {
    atomicAdd(&meta->card, px); atomicAdd(&meta->n_spans, n);
    atomicMin(&meta->first_row, r); atomicMax(&meta->last_row, r);
    atomicMin(&meta->bit_lo, lo); atomicMax(&meta->bit_hi, hi);
  }
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
// (update_objs = 0: a second leaf list over the same records only refreshes its boxes)
__global__ void k_move_leaves(ObjRec* __restrict__ objs, int4* __restrict__ leaf_box, const int* __restrict__ leaves,
                              int n_leaves, int first_obj, int last_obj, int ddx, int ddy, int update_objs, int also_obj = -1) {
  int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= n_leaves) return;
  int oi = leaves[li];
  const bool also = oi == also_obj;   // (a record outside the range that moves along: the sprite leaf of a cached group)
  if ((oi < first_obj || oi > last_obj) && !also) return;
  ObjRec& o = objs[oi];
  if (update_objs || also) { o.dx += ddx; o.dy += ddy; o.bx0 += ddx; o.bx1 += ddx; o.by0 += ddy; o.by1 += ddy; }
  leaf_box[li] = make_int4(o.bx0, o.by0, o.bx1, o.by1);
}
// ------------------------------------------------------------------------------------
// Partial-sprite cache (render.ml:1169-1242, cache.ml:328-367, 390-407).  The sprite of a cached object lives in an
// RGBA8 canvas in the object's own frame, next to two bit planes over the same box: S = its shape, V = pshape, the
// pixels whose value is in the canvas.  A frame renders only shptorender = r' - pshape (here: update ∩ S - V, a
// superset) and merges it into the canvas; everything else is served from it.  Planes are [cv_h][cv_nw] words, bit 0
// of word 0 = pixel cv_x0 of the object's frame; (dx, dy) is the object's alias offset (cache.ml:400-405).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t plane_bits32(const uint32_t* __restrict__ plane, int nw, int h, int row, int bitoff) {
  if (row < 0 || row >= h) return 0u;
  const uint32_t* r = plane + (size_t)row * nw;
  const int qw = bitoff >> 5, qb = bitoff & 31;
  const uint32_t lo = (qw >= 0 && qw < nw) ? r[qw] : 0u;
  const uint32_t hi = (qw + 1 >= 0 && qw + 1 < nw) ? r[qw + 1] : 0u;
  return qb ? ((lo >> qb) | (hi << (32 - qb))) : lo;
}
// T = update ∩ S - V as a bit-frame (frame coordinates), rows [y0, y0 + n_rows); `missing` counts its pixels
__global__ void k_sprite_todo(const uint32_t* __restrict__ U /* update bit-frame or null */, int ux0, int uy0, int ux1, int uy1,
                              const uint32_t* __restrict__ S, const uint32_t* __restrict__ V, int cv_x0, int cv_y0, int cv_nw, int cv_h,
                              int dx, int dy, int W, int nw, int y0, int n_rows, uint32_t* __restrict__ T, int* __restrict__ missing) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (w >= nw || r >= n_rows) return;
  const int y = y0 + r, x0 = 32 * w;
  uint32_t u = U ? U[(size_t)y * nw + w] : ((y >= uy0 && y <= uy1) ? interval_mask32(x0, ux0, ux1) : 0u);
  if (x0 + 31 >= W) u &= interval_mask32(x0, 0, W - 1);
  uint32_t t = 0u;
  if (u) {
    const int row = y - dy - cv_y0, bit = x0 - dx - cv_x0;
    t = u & plane_bits32(S, cv_nw, cv_h, row, bit) & ~plane_bits32(V, cv_nw, cv_h, row, bit);
  }
  T[(size_t)y * nw + w] = t;
  if (t && missing) atomicAdd(missing, __popc(t));
}
// merge the freshly rendered pixels (frame-sized canvas `tmp`, pixels of T) into the sprite's canvas and pshape
__global__ void k_sprite_store(const uint32_t* __restrict__ tmp, const uint32_t* __restrict__ T, uint32_t* __restrict__ canvas,
                               uint32_t* __restrict__ V, int cv_x0, int cv_y0, int cv_nw, int cv_h, int dx, int dy, int W, int nw, int y0, int n_rows) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (x >= W || r >= n_rows) return;
  const int y = y0 + r;
  if (!((T[(size_t)y * nw + (x >> 5)] >> (x & 31)) & 1u)) return;
  const int cx = x - dx - cv_x0, cy = y - dy - cv_y0;
  if (cx < 0 || cy < 0 || cx >= cv_nw * 32 || cy >= cv_h) return;
  canvas[(size_t)cy * (cv_nw * 32) + cx] = tmp[(size_t)y * W + x];
  atomicOr(&V[(size_t)cy * cv_nw + (cx >> 5)], 1u << (cx & 31));
}
// pixels of the shape that are not in pshape yet
__global__ void k_sprite_missing(const uint32_t* __restrict__ S, const uint32_t* __restrict__ V, size_t n, int* __restrict__ missing) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t m = S[i] & ~V[i];
  if (m) atomicAdd(missing, __popc(m));
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
                               uint32_t colour, int W, int H, int nw, uint8_t* __restrict__ alpha, uint32_t* __restrict__ unfinished) {
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
  if (lane == 0) unfinished[(size_t)y * nw + w] = t & ~f;   // pixels_for_normal_scene = shptorender' --- pixels_finished (render.ml:1105)
}
// T = S & U, and optionally a copy R of it (shptorender = r &&& u, render.ml:1281; the reading shape of a filter that reads where it writes)
__global__ void k_and_rows(const uint32_t* __restrict__ S, const uint32_t* __restrict__ U, uint32_t* __restrict__ T, uint32_t* __restrict__ R, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = S[i] & U[i];
  T[i] = t;
  if (R) R[i] = t;
}
// Which pixels of T the matte has to super-sample: all but those whose 5 x 5 neighbourhood lies in the geometry's
// minshape M = S & ~C (see apply_filter): todo = T - (box - dilate 2 2 (not M)), rows [0, h) of the planes given,
// box = columns [2, W-3] of rows [2, h-3].
__global__ void k_matte_todo(const uint32_t* __restrict__ S, const uint32_t* __restrict__ C, const uint32_t* __restrict__ T,
                             uint32_t* __restrict__ todo, int h, int nw, int W) {
  int w = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (w >= nw || r >= h) return;
  uint32_t dil = 0u;
  for (int rr = max(0, r - 2); rr <= min(h - 1, r + 2); rr++) {
    const uint32_t* s = S + (size_t)rr * nw; const uint32_t* c = C + (size_t)rr * nw;
    const uint32_t a = w > 0 ? ~(s[w - 1] & ~c[w - 1]) : 0u, b = ~(s[w] & ~c[w]), d = w + 1 < nw ? ~(s[w + 1] & ~c[w + 1]) : 0u;
    dil |= hdilate_word(a, b, d, 2);
  }
  const uint32_t box = (r >= 2 && r <= h - 3) ? interval_mask32(32 * w, 2, W - 3) : 0u;
  todo[(size_t)r * nw + w] = T[(size_t)r * nw + w] & ~(box & ~dil);
}
// blend' (render.ml:1248-1265) and the composite of the filter's sprite into the accumulator
// (render.ml:1290-1291): fb = over fb (pd_plus (dissolve Z (255 - alpha)) (dissolve Y alpha)) on T.
// Pixels a scene did not render are clear in Z / Y, which both operators treat as absent.
__device__ __forceinline__ uint32_t px_monochrome(uint32_t c) {
  const uint32_t av = ((c & 255u) + ((c >> 8) & 255u) + ((c >> 16) & 255u)) / 3u;
  return av | (av << 8) | (av << 16) | (c & 0xFF000000u);
}
// flags: 1 = the accumulator is clear on every pixel of T (nothing but filters composited so far: over clear s = s);
// 2 = Y is Colour.monochrome of Z (filters.ml:229-238, where reading scene and scene below are the same list).
// The filter's whole shape S then leaves u (the extra finish, render.ml:1120-1121, 1308).
__global__ void k_filter_blend(const uint32_t* __restrict__ T, const uint8_t* __restrict__ alpha, const uint32_t* __restrict__ Z,
                               const uint32_t* __restrict__ Y, uint32_t* __restrict__ fb, int W, int H, int nw, int flags,
                               const uint32_t* __restrict__ S, uint32_t* __restrict__ U) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t wi = (size_t)y * nw + (x >> 5);
  const uint32_t t = T[wi];
  if ((x & 31) == 0 && U) U[wi] &= ~S[wi];
  if (!((t >> (x & 31)) & 1u)) return;
  const size_t i = (size_t)y * W + x;
  const int a = alpha[i];
  const uint32_t zz = Z[i];
  const uint32_t z = px_dissolve(zz, 255 - a), yy = (flags & 2) ? px_dissolve(px_monochrome(zz), a) : (Y ? px_dissolve(Y[i], a) : 0u);
  const uint32_t res = px_plus(z, yy);
  fb[i] = (flags & 1) ? res : px_over(fb[i], res);
}
// Brush.sprite_of_brushstroke (brush.ml:176-222) for the pixels of a span set in canonical span order: the ordered
// alpha_over of every stamp that covers a pixel (207-212), then dissolve (fill x y) by that alpha (214-220).  One warp
// per row of the shape, lanes over the pixels of its spans, the stamp points walked by the whole warp.
__global__ void __launch_bounds__(128) k_brush_sprite(const int* __restrict__ row_ptr, const int2* __restrict__ spans, const int* __restrict__ px_off,
                                                      int y0, int n_rows, const int2* __restrict__ points, int n_points, int br,
                                                      const uint8_t* __restrict__ stamp, FillRec fill, uint32_t* __restrict__ out) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int y = y0 + r, w = 2 * br + 1;
  int o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    const int2 sp = spans[k];
    for (int i0 = 0; i0 < sp.y; i0 += 32) {
      const int x = sp.x + i0 + lane;
      const bool in = i0 + lane < sp.y;
      uint32_t al = 0u;
      for (int q = 0; q < n_points; q++) {
        const int2 p = points[q];
        const int ddx = x - p.x, ddy = y - p.y;
        if (ddy < -br || ddy > br) continue;   // (uniform over the warp)
        if (in && ddx >= -br && ddx <= br) al = alpha_over(al, stamp[(ddy + br) * w + (ddx + br)]);
      }
      if (in) out[o + i0 + lane] = px_dissolve(fill_lookup(fill, x, y), (int)al);
    }
    o += sp.y;
  }
}
// Alpha bytes of a canvas (cw x ch pixels with its origin at (ox, oy) of the frame) into rows [y0, y0 + h) of a dense
// byte plane `pitch` wide: the matte of a filter whose geometry is an object of its own (render.ml:1099).
__global__ void k_canvas_alpha(const uint32_t* __restrict__ canvas, int cw, int ch, int ox, int oy, int y0, int h, int pitch, uint8_t* __restrict__ op) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (x >= pitch || r >= h) return;
  const int cx = x - ox, cy = y0 + r - oy;
  op[(size_t)r * pitch + x] = (cx >= 0 && cx < cw && cy >= 0 && cy < ch) ? (uint8_t)(canvas[(size_t)cy * cw + cx] >> 24) : (uint8_t)0;
}
// Bounding box of the set bits of a bit-frame's rows [0, h): bb = {x0, y0, x1, y1} by atomic min / max (start from
// {INT_MAX, INT_MAX, INT_MIN, INT_MIN}); y counts from `ybase`.
__global__ void k_bits_bbox(const uint32_t* __restrict__ bits, int h, int nw, int ybase, int* __restrict__ bb) {
  int w = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (w >= nw || r >= h) return;
  const uint32_t v = bits[(size_t)r * nw + w];
  if (!v) return;
  atomicMin(&bb[0], 32 * w + __ffs((int)v) - 1); atomicMax(&bb[2], 32 * w + 31 - __clz((int)v));
  atomicMin(&bb[1], ybase + r); atomicMax(&bb[3], ybase + r);
}
// Brush.smear (brush.ml:286-331) on a canvas: for every smear point in order, twice over, the brush-sized block one
// step against the direction of travel is read, then blended into the block at the point by the brush's alpha
// (Colour.dissolve_between).  Each step reads what the steps before it wrote: one block walks the points; its threads
// share the pixels of the brush.  The reference's canvas is the bounding box of (the reading scene's sprite ∪ the stroke's
// shape) with a border of one pixel, and a step whose source or destination block leaves it is skipped (its exception is
// swallowed): `bb` holds the sprite's box (k_bits_bbox over the touched plane), sb the stroke's.
// X: frame-sized source canvas (W x H); C: working canvas covering [cx0, cx0 + cw) x [cy0, cy0 + ch) (clear outside the
// frame); the result goes to Y (frame-sized) wherever C overlaps the frame.
constexpr int SMEAR_THREADS = 1024;
constexpr int SMEAR_MAX_PER_THREAD = 17;   // (2 * 64 + 1)^2 / 1024
__global__ void __launch_bounds__(SMEAR_THREADS) k_smear(const uint32_t* __restrict__ X, uint32_t* __restrict__ Y, int W, int H,
                                                         uint32_t* __restrict__ C, int cx0, int cy0, int cw, int ch,
                                                         const int* __restrict__ bb, int sbx0, int sby0, int sbx1, int sby1,
                                                         const int2* __restrict__ pts, int n_pts, int odx, int ody,
                                                         const uint8_t* __restrict__ stamp, int rad,
                                                         int rx0, int ry0, int rx1, int ry1 /* the part of X that was rendered into */) {
  const int tid = threadIdx.x, bw = 2 * rad + 1, n_px = bw * bw;
  // the canvas of the reference: box of (sprite ∪ stroke shape), one pixel of border
  const int vx0 = min(bb[0], sbx0) - 1, vy0 = min(bb[1], sby0) - 1, vx1 = max(bb[2], sbx1) + 1, vy1 = max(bb[3], sby1) + 1;
  for (int i = tid; i < cw * ch; i += SMEAR_THREADS) {
    const int x = cx0 + i % cw, y = cy0 + i / cw;
    C[i] = (x >= rx0 && x <= rx1 && y >= ry0 && y <= ry1) ? X[(size_t)y * W + x] : 0u;
  }
  __syncthreads();
  uint32_t v[SMEAR_MAX_PER_THREAD];
  for (int pass = 0; pass < 2; pass++)
    for (int i = 0; i < n_pts; i++) {
      const int2 p = pts[i];
      int dx = 0, dy = 0;
      if (i) { const int2 q = pts[i - 1]; dx = p.x > q.x ? -1 : (p.x < q.x ? 1 : 0); dy = p.y > q.y ? -1 : (p.y < q.y ? 1 : 0); }   // (sic: brush.ml:268-272)
      const int px = p.x + odx, py = p.y + ody;
      // source block: centred one step away; destination block: centred at the point (both uniform over the block)
      const int sx0 = px - dx - rad, sy0 = py - dy - rad, dx0 = px - rad, dy0 = py - rad;
      if (sx0 < vx0 || sy0 < vy0 || sx0 + bw - 1 > vx1 || sy0 + bw - 1 > vy1) continue;    // Failure "subcopy"
      if (dx0 < vx0 || dy0 < vy0 || dx0 + bw - 1 > vx1 || dy0 + bw - 1 > vy1) continue;    // Failure "Brush.stamp"
#pragma unroll
      for (int k = 0; k < SMEAR_MAX_PER_THREAD; k++) {
        const int j = tid + k * SMEAR_THREADS;
        if (j < n_px) v[k] = C[(size_t)(sy0 + j / bw - cy0) * cw + (sx0 + j % bw - cx0)];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SMEAR_MAX_PER_THREAD; k++) {
        const int j = tid + k * SMEAR_THREADS;
        if (j < n_px) {
          uint32_t* d = C + (size_t)(dy0 + j / bw - cy0) * cw + (dx0 + j % bw - cx0);
          *d = px_dissolve_between(v[k], *d, stamp[j]);
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < cw * ch; i += SMEAR_THREADS) {
    const int x = cx0 + i % cw, y = cy0 + i / cw;
    if (x >= 0 && x < W && y >= 0 && y < H) Y[(size_t)y * W + x] = C[i];
  }
}
// The background list of a frame whose scene pass is finished (render.ml:1363-1365): frame = scene over background,
// for a background of plain primitives (the page, the window background): per pixel of the update that the scene
// pass left not opaque, the front-to-back fold of the primitives covering it goes under the framebuffer's colour.
struct PeerFbs { uint32_t* p[7]; };
struct BgPrims { int n; int x0[8], y0[8], x1[8], y1[8]; uint32_t col[8]; int pretrans[8]; };
__global__ void k_bg_over(uint32_t* __restrict__ fb, const uint32_t* __restrict__ u_init, int W, int nw, int ux0, int uy0, int ux1, int uy1,
                          BgPrims B, int n_peers, PeerFbs peers) {
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = uy0 + blockIdx.y;
  if (x > ux1 || x + 3 < ux0 || y > uy1) return;
  uint32_t m = 0u;
  for (int k = 0; k < 4; k++) if (x + k >= ux0 && x + k <= ux1 && x + k < W) m |= 1u << k;
  if (u_init) m &= (u_init[(size_t)y * nw + (x >> 5)] >> (x & 31)) & 15u;
  if (!m) return;
  const size_t at = (size_t)y * W + x;
  const bool vec = x + 3 < W && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(fb) & 15) == 0;
  uint32_t px[4];
  if (vec) { const uint4 v = *reinterpret_cast<const uint4*>(fb + at); px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w; }
  else for (int k = 0; k < 4; k++) px[k] = (x + k < W) ? fb[at + k] : 0u;
  bool changed = false;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (!((m >> k) & 1u) || (px[k] >> 24) == 255u) continue;
    uint32_t acc = 0u;
#pragma unroll
    for (int j = 0; j < 8; j++) {   // (constant indices: the parameter arrays stay in the constant bank)
      if (j >= B.n || (acc >> 24) == 255u) continue;
      if (x + k < B.x0[j] || x + k > B.x1[j] || y < B.y0[j] || y > B.y1[j]) continue;
      uint32_t c = B.col[j];
      if (B.pretrans[j] >= 0) c = px_dissolve(c, B.pretrans[j]);
      acc = px_over(acc, c);
    }
    if (acc != 0u) { px[k] = px_over(px[k], acc); changed = true; }
  }
  if (!changed) return;
  if (vec) {
    const uint4 v = make_uint4(px[0], px[1], px[2], px[3]);
    *reinterpret_cast<uint4*>(fb + at) = v;
    for (int p = 0; p < n_peers; p++) *reinterpret_cast<uint4*>(peers.p[p] + at) = v;
  } else
    for (int k = 0; k < 4; k++) if (x + k < W) { fb[at + k] = px[k]; for (int p = 0; p < n_peers; p++) peers.p[p][at + k] = px[k]; }
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
                            int kind /*1 unit, 2 xy*/, const int* __restrict__ taps, int total, int vertical, int cx0 = 0, int cx1 = 0x7FFFFFFF) {
  int x = cx0 + blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;   // only columns [cx0, cx1] are written
  if (x >= w || x > cx1 || y >= h) return;
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
// Pixels per row of a span set (input of the exclusive scan that gives every row its offset in span order).
__global__ void k_row_pixels(const int* __restrict__ row_ptr, const int2* __restrict__ spans, int n_rows, int* __restrict__ counts) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int n = 0;
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) n += spans[k].y;
  counts[r] = n;
}
// Polygon.polygon_sprite_edgelist's colouring (polygon.ml:733-738): every pixel of a span takes the fill at the
// span's FIRST x, dissolved by the pixel's opacity.  Thread per row, values in canonical span order.
__global__ void k_sprite_fill(const int* __restrict__ row_ptr, const int2* __restrict__ spans, const int* __restrict__ px_off,
                              int y0, int n_rows, FillRec fill, const uint8_t* __restrict__ opacity, uint32_t* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    const int2 s = spans[k];
    const uint32_t c = fill_lookup(fill, s.x, y0 + r);
    for (int i = 0; i < s.y; i++, o++) out[o] = px_dissolve(c, opacity[o]);
  }
}
// Sprite.fillshape shp fill (sprite.ml:158-175): every pixel the fill at its own coordinates (a Plain fill is one colour)
__global__ void k_sprite_fillshape(const int* __restrict__ row_ptr, const int2* __restrict__ spans, const int* __restrict__ px_off,
                                   int y0, int n_rows, FillRec fill, uint32_t* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    const int2 s = spans[k];
    for (int i = 0; i < s.y; i++, o++) out[o] = fill_lookup(fill, s.x + i, y0 + r);
  }
}
// Sprite.map_coords (fun x y c -> dissolve (fill x y) (alpha c)) (render.ml:976-981): the fill applied to an alpha matte
__global__ void k_sprite_fill_alpha(const int* __restrict__ row_ptr, const int2* __restrict__ spans, const int* __restrict__ px_off,
                                    int y0, int n_rows, FillRec fill, const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    const int2 s = spans[k];
    for (int i = 0; i < s.y; i++, o++) out[o] = px_dissolve(fill_lookup(fill, s.x + i, y0 + r), (int)(in[o] >> 24));
  }
}
// Sprite.sprite_map f (sprite.ml:358-374), f one of the colour functions of colour.ml:266-304
__global__ void k_sprite_map(int op, int arg, const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = in[i];
  uint32_t r;
  switch (op) {
    case 0: { const uint32_t av = ((c & 255u) + ((c >> 8) & 255u) + ((c >> 16) & 255u)) / 3u; r = av | (av << 8) | (av << 16) | (c & 0xFF000000u); break; }  // monochrome
    case 1: r = px_dissolve(c, arg); break;
    case 2: r = c & 0xFF0000FFu; break;   // red_channel
    case 3: r = c & 0xFF00FF00u; break;   // green_channel
    default: r = c & 0xFFFF0000u; break;  // blue_channel
  }
  out[i] = r;
}
// Scatter per-pixel values given in canonical span order into a dense canvas: thread per row.
template <class T>
__global__ void k_scatter_spans(const int* __restrict__ row_ptr, const int2* __restrict__ spans,
                                const int* __restrict__ px_off, int n_rows, int row0, int wx0, int w,
                                const T* __restrict__ in, T* __restrict__ dense) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    int2 s = spans[k];
    for (int i = 0; i < s.y; i++) dense[(size_t)(row0 + r) * w + (s.x + i - wx0)] = in[o++];
  }
}
// Gather dense per-pixel values in canonical span order: thread per row.
template <class T>
__global__ void k_gather_spans(const int* __restrict__ row_ptr, const int2* __restrict__ spans,
                               const int* __restrict__ px_off, int n_rows, int wx0, int pitch,
                               const T* __restrict__ dense, T* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int o = px_off[r];
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
    int2 s = spans[k];
    for (int i = 0; i < s.y; i++) out[o++] = dense[(size_t)r * pitch + (s.x + i - wx0)];
  }
}

// ------------------------------------------------------------------------------------
// N2: Polygon.edgelist_of_path on the device (polygon.ml:83-127, 262-287; coord.ml:47).  One thread per
// path segment: de Casteljau subdivision until Polygon.bezier_epsilon holds (curve_accuracy = 0.2), depth first,
// left half first — the pieces come out in the reference's order.  FP64 with one rounding per operation
// (__dadd_rn / __dmul_rn / __ddiv_rn / __dsqrt_rn) like OCaml.  EMIT = false counts the pieces of every segment,
// EMIT = true writes them as int32 sub-bin edges at the segment's offset (exclusive scan of the counts).
// segs: records of 9 doubles (kind 0 straight / 1 bezier, then up to 4 points).  box: x min / y min / x max / y max of
// the emitted coordinates by atomics.
// ------------------------------------------------------------------------------------
constexpr int FLATTEN_MAX_DEPTH = 40;
struct Pt2 { double x, y; };
__device__ __forceinline__ Pt2 pt_half(Pt2 a, Pt2 b) { return Pt2{__ddiv_rn(__dadd_rn(a.x, b.x), 2.), __ddiv_rn(__dadd_rn(a.y, b.y), 2.)}; }
__device__ __forceinline__ double dist_point_line(Pt2 c, Pt2 a, Pt2 b) {   // polygon.ml:83-90
  const double ex = __dadd_rn(b.x, -a.x), ey = __dadd_rn(b.y, -a.y);
  const double l = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
  const double num = __dadd_rn(__dmul_rn(__dadd_rn(a.y, -c.y), ex), -__dmul_rn(__dadd_rn(a.x, -c.x), ey));
  return __dmul_rn(fabs(__ddiv_rn(num, __dmul_rn(l, l))), l);
}
__device__ __forceinline__ bool fp_normal(double d) { return isfinite(d) && fabs(d) >= 2.2250738585072014e-308; }   // classify_float = FP_normal
__device__ __forceinline__ bool flat_enough(Pt2 p1, Pt2 p2, Pt2 p3, Pt2 p4) {   // polygon.ml:107-116
  const double d1 = dist_point_line(p2, p1, p4), d2 = dist_point_line(p3, p1, p4);
  if (fp_normal(d1) && fp_normal(d2)) return d1 < 0.2 && d2 < 0.2;
  return true;
}
__device__ __forceinline__ int sub_of_float_dev(double f) { return (int)ceil(__dadd_rn(__dmul_rn(f, 32.0), -16.0)); }   // coord.ml:47
template <bool EMIT>
__global__ void k_flatten(const double* __restrict__ segs, int n_segs, int* __restrict__ counts, const int* __restrict__ offs,
                          int4* __restrict__ edges, int* __restrict__ box, int* __restrict__ error_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_segs) return;
  const double* q = segs + 9 * (size_t)i;
  int n = 0;
  int4* out = EMIT ? edges + offs[i] : nullptr;
  int bx0 = INT32_MAX, by0 = INT32_MAX, bx1 = INT32_MIN, by1 = INT32_MIN;
  auto emit = [&](Pt2 a, Pt2 b) {
    if (EMIT) {
      const int4 e = make_int4(sub_of_float_dev(a.x), sub_of_float_dev(a.y), sub_of_float_dev(b.x), sub_of_float_dev(b.y));
      out[n] = e;
      bx0 = min(bx0, min(e.x, e.z)); bx1 = max(bx1, max(e.x, e.z)); by0 = min(by0, min(e.y, e.w)); by1 = max(by1, max(e.y, e.w));
    }
    n++;
  };
  if (q[0] == 0.) emit(Pt2{q[1], q[2]}, Pt2{q[3], q[4]});
  else {
    // explicit stack of the right halves still to visit (the left half is always taken next)
    Pt2 st[FLATTEN_MAX_DEPTH][4];
    int sp = 0;
    Pt2 p1{q[1], q[2]}, p2{q[3], q[4]}, p3{q[5], q[6]}, p4{q[7], q[8]};
    for (;;) {
      if (flat_enough(p1, p2, p3, p4)) {
        emit(p1, p4);
        if (sp == 0) break;
        sp--; p1 = st[sp][0]; p2 = st[sp][1]; p3 = st[sp][2]; p4 = st[sp][3];
        continue;
      }
      const Pt2 l2 = pt_half(p1, p2), h = pt_half(p2, p3), l3 = pt_half(l2, h), r3 = pt_half(p3, p4), r2 = pt_half(h, r3), l4 = pt_half(l3, r2);
      if (sp >= FLATTEN_MAX_DEPTH) { *error_flag = 2; break; }
      st[sp][0] = l4; st[sp][1] = r2; st[sp][2] = r3; st[sp][3] = p4; sp++;
      p2 = l2; p3 = l3; p4 = l4;
    }
  }
  if (!EMIT) counts[i] = n;
  else if (n > 0 && box) { atomicMin(&box[0], bx0); atomicMin(&box[1], by0); atomicMax(&box[2], bx1); atomicMax(&box[3], by1); }
}
