"""Scanline-band sharding across the GPUs of one box.

Every stage of the raster path is a pure function of (edge lists, row), so band k of N
owns rows [floor(k*H/N), floor((k+1)*H/N)) and renders them with no halo and no exchange;
the only exchange is the gather of the RGBA8 strips, which the library fuses into the rendering
kernels (peer stores over NVLink: coh_multi_* for one host process, coh_fb_alloc_shared / coh_fb_open_peer for one
process per GPU).  This module is the band arithmetic only; it imports nothing but numpy."""


def band_rows(height, n_bands, k):
    return k * height // n_bands, (k + 1) * height // n_bands


def all_bands(height, n_bands):
    return [band_rows(height, n_bands, k) for k in range(n_bands)]


def row_costs(edges, height, width, edge_weight=15.0, tile_weight=0.75):
    """Relative cost of every pixel row for the walker: a constant per 32-pixel tile (background fill) plus a
    weight per edge whose extended band reaches the row (scan conversion + antialiasing happen there).
    edges: int32 [n, 4] sub-pixel bins (x0, y0, x1, y1)."""
    import numpy as np

    cost = np.full(height, tile_weight * ((width + 31) // 32), dtype=np.float64)
    if len(edges):
        e = np.asarray(edges).reshape(-1, 4)
        ymin, ymax = np.minimum(e[:, 1], e[:, 3]), np.maximum(e[:, 1], e[:, 3])
        lo = np.clip((ymin - 16 + 31) >> 5, 0, height)       # the row range of k_rowedges
        hi = np.clip(((ymax + 67) >> 5) + 1, 0, height)
        d = np.zeros(height + 1, dtype=np.float64)
        np.add.at(d, lo, edge_weight)
        np.add.at(d, hi, -edge_weight)
        cost += np.cumsum(d)[:height]
    return cost


def balanced_bands(cost, n_bands, align=16):
    """Split rows into n_bands contiguous bands of about equal total cost, boundaries on multiples of `align`
    (the walker's cell height) so that no cell row is shared by two GPUs.  Returns [(y0, y1)] like all_bands."""
    import numpy as np

    height = len(cost)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for k in range(1, n_bands):
        y = int(np.searchsorted(cum, cum[-1] * k / n_bands))
        y = int(round(y / align)) * align
        lo = cuts[-1] + align                                  # every band keeps at least one cell row
        hi = height - (n_bands - k) * align
        cuts.append(max(lo, min(y, hi)) if hi >= lo else min(height, cuts[-1] + max(1, (height - cuts[-1]) // (n_bands - k + 1))))
    cuts.append(height)
    return [(cuts[k], cuts[k + 1]) for k in range(n_bands)]
