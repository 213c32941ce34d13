/* Plain C (C99) against include/coherence_b200.h: calls EVERY exported entry point with plain host buffers, and every
 * error path that is reachable without breaking the device.  Prints "ok <name>" per check and "ABI-EVERY-SYMBOL PASS n"
 * at the end; any unexpected status prints "FAIL ..." and exits 1.  Without a CUDA device it checks that coh_init
 * fails loudly (there is no CPU path) and exits 0.  tests/test_abi_c.py builds and runs it. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "coherence_b200.h"

static coh_ctx* C = NULL;
static int n_ok = 0;
#define OK(call) do { if (getenv("ABI_TRACE")) printf("> %s\n", #call); if ((call) != 0) { printf("FAIL %s: %s\n", #call, coh_last_error(C)); exit(1); } n_ok++; } while (0)
#define ERR(call, what) do { if ((call) == 0) { printf("FAIL expected an error: %s\n", #call); exit(1); } \
  if (!strstr(coh_last_error(C), what)) { printf("FAIL wrong message for %s: %s\n", #call, coh_last_error(C)); exit(1); } n_ok++; } while (0)
#define CHECK(cond) do { if (!(cond)) { printf("FAIL %s (line %d)\n", #cond, __LINE__); exit(1); } n_ok++; } while (0)

static int sub_of_float(double f) { return (int)ceil(f * 32.0 - 16.0); }
static void tri_edges(int32_t* e, double ax, double ay, double bx, double by, double cx, double cy) {
  double p[4][2] = {{ax, ay}, {bx, by}, {cx, cy}, {ax, ay}};
  for (int i = 0; i < 3; i++) { e[4 * i] = sub_of_float(p[i][0]); e[4 * i + 1] = sub_of_float(p[i][1]); e[4 * i + 2] = sub_of_float(p[i + 1][0]); e[4 * i + 3] = sub_of_float(p[i + 1][1]); }
}
static coh_object blank(int kind) { coh_object o; memset(&o, 0, sizeof o); o.kind = kind; o.pretrans = -1; o.id = -1; return o; }

int main(int argc, char** argv) {
  (void)argc; (void)argv;
  setvbuf(stdout, NULL, _IONBF, 0);
  if (coh_init(-1, &C) != 0) {
    const char* m = coh_last_error(NULL);
    if (!strstr(m, "no CPU fallback")) { printf("FAIL init message: %s\n", m); return 1; }
    printf("no device: %s\nABI-EVERY-SYMBOL NO-DEVICE\n", m);
    return 0;
  }
  char name[128];
  OK(coh_device_name(C, name, sizeof name));
  CHECK(coh_stream(C) != NULL);
  CHECK(coh_launch_count(C) >= 0);
  OK(coh_set_option(C, "walk_h", 0));
  ERR(coh_set_option(C, "walk_h", 3), "walk_h");
  ERR(coh_set_option(C, "no_such_option", 1), "unknown option");
  OK(coh_set_timing(C, 1));
  /* colour codec: Colour.clear = 0x43800000, white = 0x7F9FFFFF */
  CHECK(coh_colour_of_rgba8(0u) == 0x43800000); CHECK(coh_colour_of_rgba8(0xFFFFFFFFu) == 0x7F9FFFFF);
  CHECK(coh_rgba8_of_colour(0x7F9FFFFF) == 0xFFFFFFFFu);
  /* Polygon */
  int32_t e[12]; tri_edges(e, 20.3, 20.1, 120.7, 30.2, 60.2, 150.9);
  coh_shape_t s = 0, m = 0, mx = 0, t = 0, u = 0;
  OK(coh_shapeminshape_of_edgelist(C, e, 3, COH_NONZERO, &s, &m));
  ERR(coh_shapeminshape_of_edgelist(C, e, 3, 7, &t, &u), "winding");
  OK(coh_shapeminshape_of_edgelist(C, e, 0, COH_NONZERO, &t, &u)); CHECK(t == 0 && u == 0);   /* NullShape, NullShape */
  int64_t card = 0, n = 0, sz = 0;
  OK(coh_shape_card(C, s, &card)); CHECK(card > 0);
  OK(coh_shape_difference(C, s, m, &mx));
  OK(coh_shape_card(C, mx, &card));
  uint8_t* op = (uint8_t*)malloc((size_t)card);
  OK(coh_polygon_opacity(C, e, 3, COH_NONZERO, mx, op, card, &n)); CHECK(n == card);
  ERR(coh_polygon_opacity(C, e, 3, COH_NONZERO, mx, op, card - 1, &n), "buffer too small");
  coh_object fill = blank(COH_OBJ_PATH); fill.fill_kind = COH_FILL_PLAIN; fill.colour0 = 0xFF2030C8u;
  uint32_t* px = (uint32_t*)malloc(4 * (size_t)200 * 160);   /* large enough for every sprite read below */
  OK(coh_polygon_sprite(C, &fill, e, 3, COH_NONZERO, mx, px, card, &n)); CHECK(n == card);
  ERR(coh_polygon_sprite(C, &fill, e, 3, COH_NONZERO, mx, px, 1, &n), "buffer too small");
  fill.fill_kind = 9; ERR(coh_polygon_sprite(C, &fill, e, 3, COH_NONZERO, mx, px, card, &n), "fill kind"); fill.fill_kind = 0;
  /* Sprite span-set algebra */
  coh_shape_t bx = 0, un = 0, in = 0, tr = 0, bl = 0, er = 0, imp = 0;
  OK(coh_shape_box(C, 10, 10, 50, 40, &bx));
  ERR(coh_shape_box(C, 0, 0, -1, 3, &t), "negative");
  OK(coh_shape_box(C, 0, 0, 0, 0, &t)); CHECK(t == 0);
  OK(coh_shape_union(C, s, bx, &un)); OK(coh_shape_intersection(C, s, bx, &in));
  OK(coh_shape_translate(C, bx, 3, -4, &tr)); OK(coh_shape_bloat(C, bx, 2, 1, &bl)); OK(coh_shape_erode(C, bx, 2, 1, &er));
  int32_t box[4], isnull = 0;
  OK(coh_shape_bounds(C, bl, box, &isnull)); CHECK(!isnull && box[0] == 8 && box[1] == 9 && box[2] == 61 && box[3] == 50);
  OK(coh_shape_bounds(C, 0, box, &isnull)); CHECK(isnull);
  OK(coh_shape_export_size(C, bx, &sz)); CHECK(sz == 40 * 4);
  int32_t* flat = (int32_t*)malloc(4 * (size_t)sz);
  OK(coh_shape_export(C, bx, flat, sz, &n)); CHECK(n == sz && flat[0] == 10 && flat[1] == 1 && flat[2] == 10 && flat[3] == 50);
  ERR(coh_shape_export(C, bx, flat, sz - 1, &n), "too small");
  OK(coh_shape_import(C, flat, sz, &imp));
  flat[3] = 0; ERR(coh_shape_import(C, flat, sz, &t), "malformed shape"); flat[3] = 50;   /* empty span: not canonical (sprite.ml:201-239) */
  /* whole-sprite operations */
  {
    int32_t yes = 0;
    OK(coh_shape_intersects(C, s, bx, &yes)); CHECK(yes == 1);
    OK(coh_shape_card(C, bx, &card));
    uint32_t* whole = (uint32_t*)malloc(4 * (size_t)card);
    OK(coh_sprite_fillshape(C, bx, &fill, whole, card, &n)); CHECK(n == card && whole[0] == 0xFF2030C8u);
    ERR(coh_sprite_fillshape(C, bx, &fill, whole, 2, &n), "buffer too small");
    int64_t ecard = 0; OK(coh_shape_card(C, er, &ecard));
    OK(coh_sprite_portion(C, bx, whole, er, px, ecard, &n)); CHECK(n == ecard && px[0] == 0xFF2030C8u);
    ERR(coh_sprite_portion(C, er, whole, bx, px, card, &n), "bad input");   /* the shape is not inside the sprite's */
    OK(coh_sprite_map(C, COH_MAP_DISSOLVE, 128, whole, card, px)); CHECK((px[0] >> 24) == 128);
    OK(coh_sprite_map(C, COH_MAP_RED_CHANNEL, 0, whole, card, px)); CHECK(px[0] == 0xFF0000C8u);
    ERR(coh_sprite_map(C, 9, 0, whole, card, px), "unknown colour function");
    ERR(coh_sprite_map(C, COH_MAP_DISSOLVE, 300, whole, card, px), "delta");
    OK(coh_sprite_map_coords_fill(C, bx, &fill, whole, px, card, &n)); CHECK(n == card && px[0] == 0xFF2030C8u);
    free(whole);
  }
  /* Convolve */
  OK(coh_shape_card(C, bx, &card));
  uint32_t* spr = (uint32_t*)malloc(4 * (size_t)card);
  for (int64_t i = 0; i < card; i++) spr[i] = 0xFF0000FFu;
  coh_shape_t cs = 0; int64_t bcard = 0;
  OK(coh_shape_card(C, bl, &bcard));
  uint32_t* cout = (uint32_t*)malloc(4 * (size_t)(bcard + 4096));
  OK(coh_convolve_sprite(C, COH_CONV_UNIT, 1, bx, spr, &cs, cout, bcard + 4096, &n)); CHECK(cs != 0 && n > card);
  ERR(coh_convolve_sprite(C, 7, 1, bx, spr, &t, cout, bcard, &n), "Invalid_argument");
  ERR(coh_convolve_sprite(C, COH_CONV_UNIT, 0, bx, spr, &t, cout, bcard, &n), "Invalid_argument");
  /* Cache */
  int64_t st[4]; int32_t found = 0;
  OK(coh_cache_clear(C)); OK(coh_cache_configure(C, 1, 32 << 20));
  OK(coh_cache_addshape(C, 41, s, m));
  OK(coh_cache_getshape(C, 41, &t, &u, &found)); CHECK(found && t && u); OK(coh_shape_free(C, t)); OK(coh_shape_free(C, u));
  OK(coh_cache_getshape(C, 42, &t, &u, &found)); CHECK(!found);
  OK(coh_cache_addtranslation(C, 42, 41, 5, 6));
  OK(coh_cache_getshape(C, 42, &t, &u, &found)); CHECK(found); OK(coh_shape_free(C, t)); OK(coh_shape_free(C, u));
  OK(coh_cache_stats(C, st)); CHECK(st[0] == 2 && st[1] == 1 && st[3] >= 1);
  coh_shape_t dirty = 0;
  OK(coh_dirty_region(C, s, m, tr, er, bl, 1, &dirty)); OK(coh_shape_free(C, dirty));
  OK(coh_dirty_region(C, s, 0, tr, 0, bl, 0, &dirty));
  /* Render */
  const int W = 200, H = 160;
  coh_object objs[5];
  objs[0] = blank(COH_OBJ_GROUP_BEGIN); objs[0].id = 7;
  objs[1] = blank(COH_OBJ_PATH); objs[1].first = 0; objs[1].count = 3; objs[1].colour0 = 0xB4141487u;   /* premultiplied, alpha 180 */
  objs[2] = blank(COH_OBJ_GROUP_END);
  objs[3] = blank(COH_OBJ_PRIMITIVE); objs[3].colour0 = 0xFF00FF00u; objs[3].prim[0] = 100; objs[3].prim[1] = 20; objs[3].prim[2] = 150; objs[3].prim[3] = 60; objs[3].id = 9;
  objs[4] = blank(COH_OBJ_PRIMITIVE); objs[4].colour0 = 0xFFD3D3D3u; objs[4].prim[2] = W; objs[4].prim[3] = H;
  coh_scene_t sc = 0, bad = 0;
  ERR(coh_render_frame(C, 0, 0, 0, W, H, 0), "coh_fb_configure");
  ERR(coh_fb_configure(C, 0, 10, 0, 10), "bad size"); ERR(coh_fb_configure(C, W, H, 10, H + 1), "bad band");
  OK(coh_fb_configure(C, W, H, 0, H));
  OK(coh_scene_create(C, objs, 5, 1, e, 3, NULL, 0, &sc));
  ERR(coh_scene_create(C, objs, 2, 0, e, 3, NULL, 0, &bad), "unterminated");
  objs[1].count = 9; ERR(coh_scene_create(C, objs, 5, 1, e, 3, NULL, 0, &bad), "out of bounds"); objs[1].count = 3;
  objs[1].kind = 77; ERR(coh_scene_create(C, objs, 5, 1, e, 3, NULL, 0, &bad), "unknown object kind"); objs[1].kind = COH_OBJ_PATH;
  OK(coh_render_frame(C, sc, 0, 0, W, H, COH_RENDER_RECORD_U));
  ERR(coh_render_frame(C, sc, 0, 0, -1, 5, 0), "negative");
  ERR(coh_render_frame(C, 0, 0, 0, W, H, 0), "null scene");
  OK(coh_sync(C));
  coh_shape_t unc = 0; OK(coh_render_uncovered(C, &unc)); CHECK(unc != 0);   /* u after the scene pass: the update minus what the triangle's interior and the green rectangle made opaque */
  OK(coh_shape_card(C, unc, &card)); CHECK(card > 0 && card < (int64_t)W * H); OK(coh_shape_free(C, unc));
  ERR(coh_shape_bloat(C, bx, -1, 0, &t), "negative radius");
  uint32_t* img = (uint32_t*)malloc(4 * (size_t)W * H);
  uint8_t* rgb = (uint8_t*)malloc(3 * (size_t)W * H);
  OK(coh_fb_read_rgba(C, 0, 0, W, H, (uint8_t*)img)); CHECK(img[0] == 0xFFD3D3D3u && img[40 * W + 140] == 0xFF00FF00u);
  ERR(coh_fb_read_rgba(C, 0, 0, W + 1, H, (uint8_t*)img), "outside");
  OK(coh_fb_read_rgb888(C, 0, 0, W, H, rgb)); CHECK(rgb[0] == 0xD3 && rgb[3 * (40 * W + 140) + 1] == 0xFF);
  ERR(coh_fb_read_rgb888(C, -1, 0, 4, 4, rgb), "outside");
  /* N4 / N1: the socket format and the RefreshWindow message of a dirty rectangle */
  {
    const int32_t kinds[4] = {COH_WIRE_TUPLE, COH_WIRE_STRING, COH_WIRE_INT, COH_WIRE_BOOL};
    const int64_t vals[4] = {3, 8, 7, 1}, offs[4] = {0, 0, 0, 0};
    uint8_t wm[64]; int32_t k2[8], nt = 0; int64_t v2[8], o2[8], taken = -1;
    int64_t wl = coh_host_wire_marshal(kinds, vals, offs, 4, (const uint8_t*)"MouseNow", wm, sizeof wm);
    CHECK(wl == 4 + 5 + 13 + 5 + 2 && wm[3] == wl - 4 && wm[4] == COH_WIRE_TUPLE && wm[8] == wl - 9 && !memcmp(wm + 14, "MouseNow", 8));
    CHECK(coh_host_wire_marshal(kinds, vals, offs, 3, (const uint8_t*)"MouseNow", wm, sizeof wm) == -1);   /* a member missing */
    CHECK(coh_host_wire_unmarshal(wm, wl - 1, k2, v2, o2, 8, &nt, &taken) == 0 && taken == 0);            /* None: incomplete */
    CHECK(coh_host_wire_unmarshal(wm, wl, k2, v2, o2, 8, &nt, &taken) == 0 && taken == wl && nt == 4 && k2[0] == COH_WIRE_TUPLE && v2[0] == 3 &&
          k2[1] == COH_WIRE_STRING && v2[1] == 8 && o2[1] == 14 && v2[2] == 7 && k2[3] == COH_WIRE_BOOL && v2[3] == 1);
    wm[9] = 9; CHECK(coh_host_wire_unmarshal(wm, wl, k2, v2, o2, 8, &nt, &taken) == -1);                    /* Invalid_data */
    uint8_t hdr[64]; int32_t hl = 0;
    CHECK(coh_host_wire_refresh_window(1, 130, 30, 149, 49, hdr, &hl) == 57 + 20 * 20 * 3 && hl == 57 && !memcmp(hdr + 14, "RefreshWindow", 13));
    CHECK(coh_host_wire_refresh_window(1, 130, 30, 130, 49, hdr, &hl) == 0 && coh_host_wire_refresh_window(1, 131, 30, 130, 49, hdr, &hl) == -1);
    uint8_t* msg = (uint8_t*)malloc(57 + 20 * 20 * 3); int64_t ml = -1;
    OK(coh_wire_refresh_window(C, 1, 130, 30, 149, 49, NULL, 0, &ml)); CHECK(ml == 57 + 20 * 20 * 3);
    OK(coh_wire_refresh_window(C, 1, 130, 30, 149, 49, msg, ml, &ml));
    CHECK(!memcmp(msg, hdr, 57) && msg[4] == COH_WIRE_TUPLE && !memcmp(msg + 57, rgb + 3 * (30 * W + 130), 60) && !memcmp(msg + 57 + 60 * 10, rgb + 3 * (40 * W + 130), 60));
    OK(coh_wire_refresh_window(C, 1, 130, 30, 130, 49, msg, ml, &ml)); CHECK(ml == 0);
    ERR(coh_wire_refresh_window(C, 1, 130, 30, W, 49, msg, 1 << 20, &ml), "outside");
    ERR(coh_wire_refresh_window(C, 1, 131, 30, 130, 49, msg, 1 << 20, &ml), "not a rectangle");
    free(msg);
  }
  OK(coh_fb_read_rgba_async(C, 0, 0, W, H, (uint8_t*)img)); OK(coh_fb_read_wait(C));
  OK(coh_shape_card(C, bx, &card));
  OK(coh_fb_read_sprite(C, bx, px, card, &n)); CHECK(n == card);
  ERR(coh_fb_read_sprite(C, bx, px, 3, &n), "too small");
  OK(coh_render_frame_shape(C, sc, bx, 0));
  coh_shape_t os = 0, om = 0;
  OK(coh_scene_object_shape(C, sc, 0, &os, &om)); CHECK(os != 0 && om == 0);   /* a Group: minshape null (render.ml:494) */
  ERR(coh_scene_object_shape(C, sc, 99, &t, &u), "no such object");
  OK(coh_scene_translate_object(C, sc, 3, 4, 5));
  ERR(coh_scene_translate_object(C, sc, 2, 1, 1), "no such object");   /* a GROUP_END is not an object */
  int32_t bb[4];
  OK(coh_scene_drag_object(C, sc, 0, 3, 2, 0, bb)); CHECK(bb[2] >= bb[0] && bb[3] >= bb[1]);
  coh_shape_t df = 0;
  OK(coh_dirty_filter(C, sc, -1, dirty, &df));
  int64_t sst[4];
  OK(coh_cache_sprite_stats(C, sc, sst)); CHECK(sst[3] == 0);   /* the group has a single member: nothing worth a sprite */
  ERR(coh_cache_sprite_stats(C, 0, sst), "null scene");
  double walk = 0, bin = 0; int64_t frames = 0;
  OK(coh_get_timing(C, &walk, &bin, &frames)); CHECK(frames >= 1); OK(coh_set_timing(C, 0));
  /* caller-owned framebuffer, peers (a second pointer into the same buffer stands in for a peer GPU) */
  void* own = coh_fb_device_ptr(C); CHECK(own != NULL);
  void* peers[1] = {own};
  OK(coh_fb_set_peers(C, 1, peers)); OK(coh_render_frame(C, sc, 0, 0, W, H, 0)); OK(coh_fb_set_peers(C, 0, NULL));
  ERR(coh_fb_set_peers(C, 9, peers), "at most 7");
  OK(coh_fb_attach(C, NULL));
  int64_t mem = 0; OK(coh_mem_in_use(C, &mem)); CHECK(mem > 0);
  OK(coh_set_stream(C, coh_stream(C)));
  /* one process per GPU: a framebuffer allocated for export; mapping it back into the SAME process is refused by CUDA */
  uint8_t ipc[64]; void* mapped = NULL;
  OK(coh_fb_alloc_shared(C, ipc));
  ERR(coh_fb_open_peer(C, ipc, &mapped), "cudaIpcOpenMemHandle");
  OK(coh_render_frame(C, sc, 0, 0, W, H, 0)); OK(coh_fb_read_rgba(C, 0, 0, W, H, (uint8_t*)img)); CHECK(img[0] == 0xFFD3D3D3u);
  /* frame signals: counters behind the pixels of a shared framebuffer (here the context signals itself) */
  {
    void* self_fb[1] = {coh_fb_device_ptr(C)}; int32_t slots[2] = {3, 3}; int dummy;
    void* stranger[1] = {(void*)&dummy};
    OK(coh_frame_signal(C, 1, self_fb, 3, 7)); OK(coh_frame_wait(C, 2, slots, 7)); OK(coh_frame_wait(C, 1, slots, 6)); OK(coh_sync(C));
    ERR(coh_frame_signal(C, 1, stranger, 3, 8), "coh_fb_open_peer");
    ERR(coh_frame_signal(C, 1, self_fb, 99, 8), "slot");
    slots[0] = -1; ERR(coh_frame_wait(C, 1, slots, 7), "slot");
  }
  OK(coh_fb_attach(C, NULL));
  { int32_t s0 = 0; ERR(coh_frame_wait(C, 1, &s0, 1), "shared"); }
  /* one process, several devices (here: as many as are visible, at most 2) */
  {
    coh_multi* M = NULL;
    int nd = 0;
    if (coh_multi_init(99, NULL, &M) == 0) { printf("FAIL coh_multi_init accepted 99 devices\n"); return 1; }
    CHECK(strstr(coh_multi_last_error(NULL), "devices") != NULL);
    for (nd = 2; nd >= 1; nd--) if (coh_multi_init(nd, NULL, &M) == 0) break;
    CHECK(M != NULL && coh_multi_device_count(M) == nd && coh_multi_ctx(M, 0) != NULL && coh_multi_ctx(M, nd) == NULL);
    if (coh_multi_configure(M, W, H, NULL)) { printf("FAIL coh_multi_configure: %s\n", coh_multi_last_error(M)); return 1; }
    int32_t badcuts[3] = {0, 50, 90};
    CHECK(coh_multi_configure(M, W, H, badcuts) != 0 || nd != 2);
    if (coh_multi_configure(M, W, H, NULL)) return 1;
    coh_scene_t msc = 0;
    if (coh_multi_scene_create(M, objs, 5, 1, e, 3, NULL, 0, &msc)) { printf("FAIL coh_multi_scene_create: %s\n", coh_multi_last_error(M)); return 1; }
    if (coh_multi_render_frame(M, msc, 0, 0, W, H, 0) || coh_multi_sync(M)) { printf("FAIL coh_multi_render_frame: %s\n", coh_multi_last_error(M)); return 1; }
    uint32_t* img2 = (uint32_t*)malloc(4 * (size_t)W * H);
    if (coh_multi_fb_read_rgba(M, 0, 0, W, H, (uint8_t*)img2)) return 1;
    coh_scene_t sc2 = 0;   /* (sc has had objects moved since) */
    OK(coh_scene_create(C, objs, 5, 1, e, 3, NULL, 0, &sc2));
    OK(coh_fb_configure(C, W, H, 0, H)); OK(coh_render_frame(C, sc2, 0, 0, W, H, 0)); OK(coh_fb_read_rgba(C, 0, 0, W, H, (uint8_t*)img));
    OK(coh_scene_free(C, sc2));
    CHECK(memcmp(img, img2, 4 * (size_t)W * H) == 0);   /* the banded frame of nd devices = the single-device frame */
    if (coh_multi_fb_read_rgb888(M, 0, 0, W, H, rgb)) return 1;
    CHECK(rgb[0] == 0xD3);
    if (coh_multi_scene_translate_object(M, msc, 3, 2, 2) || coh_multi_render_frame(M, msc, 0, 0, W, H, 0) || coh_multi_sync(M)) return 1;
    CHECK(coh_multi_render_frame(M, 0, 0, 0, W, H, 0) != 0 && strstr(coh_multi_last_error(M), "null scene") != NULL);
    if (coh_multi_scene_free(M, msc) || coh_multi_shutdown(M)) return 1;
    printf("ok coh_multi on %d device(s)\n", nd);
    free(img2);
  }
  /* host-side geometry */
  double segs[9] = {0, 10.0, 10.0, 90.0, 20.0, 0, 0, 0, 0};
  int32_t he[8]; CHECK(coh_host_edgelist_of_subpath(segs, 1, he, 2) == 1 && he[0] == sub_of_float(10.0));
  int32_t hp[64]; CHECK(coh_host_brush_points(segs, 1, 4.0, hp, 32) > 0);
  int32_t hs[256]; CHECK(coh_host_smear_points(segs, 1, hs, 128) > 30 && hs[0] == 10 && hs[1] == 10);
  /* brush strokes outside a scene */
  {
    coh_object bo; memset(&bo, 0, sizeof bo);
    bo.kind = COH_OBJ_BRUSH; bo.brush_radius = 4.0; bo.brush_opacity = 0.9; bo.pretrans = -1; bo.id = -1; bo.colour0 = 0xFF2080F0u;
    int32_t np = (int32_t)coh_host_brush_points(segs, 1, 4.0, hp, 32); if (np > 32) np = 32;   /* (the count is returned whatever cap is) */
    coh_shape_t bsh = 0, so = 0; int64_t bc = 0, bn = 0;
    OK(coh_brush_shape(C, &bo, hp, np, &bsh)); OK(coh_shape_card(C, bsh, &bc)); CHECK(bc > 100);
    uint32_t* bpx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)bc * 2 + 16);
    OK(coh_brush_sprite(C, &bo, hp, np, bsh, bpx, bc, &bn)); CHECK(bn == bc);
    ERR(coh_brush_sprite(C, &bo, hp, np, bsh, bpx, 1, &bn), "buffer too small");
    int32_t ns = (int32_t)coh_host_smear_points(segs, 1, hs, 128); if (ns > 128) ns = 128;
    OK(coh_brush_smear(C, bsh, bpx, &bo, hp, np, hs, ns, &so, bpx, bc * 2, &bn)); CHECK(so != 0 && bn == bc);
    bo.winding = COH_BRUSH_DUMMY;
    ERR(coh_brush_smear(C, bsh, bpx, &bo, hp, np, hs, ns, &so, bpx, bc * 2, &bn), "dummy brush");
    bo.winding = 7; ERR(coh_brush_shape(C, &bo, hp, np, &bsh), "brush kind");
    free(bpx); OK(coh_shape_free(C, bsh)); OK(coh_shape_free(C, so));
  }
  /* N2: the same flattening on the device */
  {
    double tri[27] = {0, 10.0, 10.0, 90.0, 20.0, 0, 0, 0, 0,   1, 90.0, 20.0, 120.0, 60.0, 40.0, 90.0, 30.0, 70.0,   0, 30.0, 70.0, 10.0, 10.0, 0, 0, 0, 0};
    int32_t de[4 * 64], he2[4 * 64]; int64_t dn = 0;
    OK(coh_edgelist_of_path(C, tri, 3, de, 64, &dn));
    int64_t hn = coh_host_edgelist_of_subpath(tri, 3, he2, 64);
    CHECK(dn == hn && dn > 3 && memcmp(de, he2, sizeof(int32_t) * 4 * (size_t)dn) == 0);
    coh_shape_t ps = 0, pm = 0, es = 0, em = 0;
    OK(coh_shapeminshape_of_path(C, tri, 3, COH_NONZERO, &ps, &pm));
    OK(coh_shapeminshape_of_edgelist(C, he2, (int32_t)hn, COH_NONZERO, &es, &em));
    int64_t c1 = 0, c2 = 0; OK(coh_shape_card(C, ps, &c1)); OK(coh_shape_card(C, es, &c2)); CHECK(c1 == c2 && c1 > 1000);
    ERR(coh_shapeminshape_of_path(C, tri, 3, 9, &ps, &pm), "winding");
    OK(coh_shape_free(C, ps)); OK(coh_shape_free(C, pm)); OK(coh_shape_free(C, es)); OK(coh_shape_free(C, em));
  }
  /* N2: the stroker — outline on the host, flattening and scan conversion on the device */
  {
    double path[27] = {0, 20.0, 20.0, 120.0, 20.0, 0, 0, 0, 0,   0, 120.0, 20.0, 120.0, 90.0, 0, 0, 0, 0,   1, 120.0, 90.0, 100.0, 130.0, 60.0, 130.0, 40.0, 90.0};
    int32_t counts[1] = {3}, oc[4], om_ = 0, ow = -1;
    coh_strokespec sp; memset(&sp, 0, sizeof sp);
    sp.startcap = COH_CAP_ROUND; sp.join = COH_JOIN_MITRED; sp.endcap = COH_CAP_PROJECTING; sp.mitrelimit = 10.0; sp.linewidth = 8.0;
    int64_t on = coh_host_strokepath(&sp, path, counts, 1, NULL, 0, oc, 4, &om_, &ow);
    CHECK(on > 12 && om_ == 1 && oc[0] == on && ow == COH_EVENODD);
    int32_t sb[4]; CHECK(coh_host_bounds_stroke(&sp, path, counts, 1, sb) == 0 && sb[0] <= 20 - 80 && sb[1] >= 120 + 80 && sb[2] <= 20 - 80 && sb[3] >= 120 + 80);   /* mitre limit 10 x width 8 */
    CHECK(coh_host_bounds_stroke(&sp, path, counts, 0, sb) == -1);
    double* outl = (double*)malloc(sizeof(double) * 9 * (size_t)on);
    CHECK(coh_host_strokepath(&sp, path, counts, 1, outl, on, oc, 4, &om_, &ow) == on && outl[0] == 1.0);   /* a round cap comes first */
    int32_t* se = (int32_t*)malloc(sizeof(int32_t) * 4 * 4096); int64_t sn = 0; int32_t sw = -1;
    OK(coh_strokepath(C, &sp, path, counts, 1, se, 4096, &sn, &sw)); CHECK(sn >= on && sn <= 4096 && sw == COH_EVENODD);
    for (int64_t i = 1; i < sn; i++) CHECK((se[4 * i + 1] > se[4 * i + 3] ? se[4 * i + 1] : se[4 * i + 3]) <= (se[4 * i - 3] > se[4 * i - 1] ? se[4 * i - 3] : se[4 * i - 1]));   /* sort_edgelist_maxy_rev */
    coh_shape_t ss = 0, sm = 0, es2 = 0, em2 = 0; int64_t c1 = 0, c2 = 0;
    OK(coh_shapeminshape_of_stroke(C, &sp, path, counts, 1, &ss, &sm));
    OK(coh_shapeminshape_of_edgelist(C, se, (int32_t)sn, sw, &es2, &em2));
    OK(coh_shape_card(C, ss, &c1)); OK(coh_shape_card(C, es2, &c2)); CHECK(c1 == c2 && c1 > 8 * 200);
    sp.join = 9; ERR(coh_strokepath(C, &sp, path, counts, 1, se, 4096, &sn, &sw), "cap or join");
    OK(coh_shape_free(C, ss)); OK(coh_shape_free(C, sm)); OK(coh_shape_free(C, es2)); OK(coh_shape_free(C, em2));
    free(outl); free(se);
  }
  /* release */
  coh_shape_t all[] = {s, m, mx, bx, un, in, tr, bl, er, imp, cs, dirty, os, om, df};
  for (unsigned i = 0; i < sizeof all / sizeof all[0]; i++) OK(coh_shape_free(C, all[i]));
  OK(coh_scene_free(C, sc)); OK(coh_cache_clear(C));
  OK(coh_shutdown(C));
  printf("ABI-EVERY-SYMBOL PASS %d\n", n_ok);
  return 0;
}
