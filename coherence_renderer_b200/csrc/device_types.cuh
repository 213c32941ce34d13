// device_types.cuh — records shared by the kernels and the C-ABI host glue.
#pragma once
#include <stdint.h>
#include "raster_core.cuh"

namespace coh {

constexpr int TILE_W = 32;       // pixels per tile word (one warp lane per pixel)
#ifndef COH_CELL_H
#define COH_CELL_H 16
#endif
constexpr int CELL_H = COH_CELL_H;  // rows per cell (power of two <= 32)
constexpr int MAX_DEPTH = 6;     // nested Group depth (incl. the implicit scene / background groups)
constexpr int AA_WORDS = 17;     // (32 + 2) * 16 scaled columns of one tile = 544 bits

enum { K_PATH = 0, K_PRIM = 1, K_GROUP = 2, K_BRUSH = 4, K_CONV = 5, K_CPG = 6 };
enum { OF_ROOT_SCENE = 1, OF_ROOT_BACKGROUND = 2,
       OF_OCCLUDES = 4 };  // leaf with an opaque plain fill and no dissolve on the way to the root: its minshape hides what lies behind

// One record per leaf object or group, device resident (96 bytes + fill).
struct ObjRec {
  int kind;
  int winding;      // shape winding rule (polygon.ml:588-591)
  int aa_winding;   // sprite winding rule (render.ml:1018 uses EvenOdd for stroked paths)
  int first, count; // edges (PATH) or stamp points (BRUSH)
  int pretrans;     // -1 or 0..255
  int dx, dy;       // alias translation (pixels)
  int bx0, by0, bx1, by1;  // conservative pixel bbox of the shape in DEVICE space (alias applied)
  int prim[4];      // PRIM: inclusive box in the object's own frame
  int depth;        // number of enclosing groups (>= 1: implicit root group)
  int anc[MAX_DEPTH];  // enclosing group object indices, outermost first
  int flags;
  int brush_r;      // BRUSH: (w-1)/2
  int stamp_off;    // BRUSH: offset of the (2r+1)^2 alpha stamp in the stamp pool
  int ry0, ry1;     // PATH: pixel rows (object frame) that have a candidate edge list
  int row_base;     // PATH: first slot of this object in the row-edge CSR
  union {
    struct {
      // CONV (Convolved (kernel, Basic (fill, Path))): pre-convolved canvas of the object, object frame
      int cv_x0, cv_y0; // pixel of bit 0 / first row of the canvas
      int cv_nw, cv_h;  // 32-pixel words per row, rows
      int cv_bits;      // offset (words) of the shape bit-rows in the scene's conv bit pool; minshape rows follow
      int cv_px;        // offset (pixels) of the convolved RGBA8 canvas in the scene's conv pixel pool
    };
    struct {
      // CPG (op, Path a, Path b): operand a uses first / count / winding / ry0 / ry1 / row_base
      int b_first, b_count;  // edges of operand b
      int b_ry0, b_ry1;      // rows of operand b that have a candidate edge list
      int b_row_base;        // first slot of operand b in the row-edge CSR
      int b_opw;             // op (COH_CPG_*) | winding rule of b << 8
    };
    struct {
      // BRUSH: grid of 32 x CELL_H pixel cells (object frame) over the stroke's box; per cell the range of
      // stamp indices (list order) whose footprint reaches it
      int bc_x0, bc_y0;      // cell column / row of the box's top-left corner: floor(x / 32), floor(y / CELL_H)
      int bc_nx, bc_ny;      // cells per row, rows of cells
      int bc_base;           // first slot of this stroke in the scene's brush range table
      int bc_pad;
    };
  };
  int pad;
  FillRec fill;
};

struct Frame {
  int W, H;            // framebuffer size in pixels
  int band_y0, band_y1;  // rows rendered by this context
  int tiles_x;         // ceil(W / 32)
  int cells_y;         // ceil(H / CELL_H)
  int ctx0, cntx;      // cell grid of the current pass: first tile column and number of tile columns (an update box
                       // usually spans few of the frame's columns); cell = row * cntx + (tile - ctx0)
};

}  // namespace coh
