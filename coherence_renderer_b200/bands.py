"""Scanline-band sharding across the GPUs of one box.

Every stage of the raster path is a pure function of (edge lists, row), so band k of N
owns rows [floor(k*H/N), floor((k+1)*H/N)) and renders them with no halo and no exchange;
the only collective is the gather of the RGBA8 strips into the output framebuffer.
torch.distributed is plumbing (process group + NCCL/gloo transport)."""


def band_rows(height, n_bands, k):
    return k * height // n_bands, (k + 1) * height // n_bands


def all_bands(height, n_bands):
    return [band_rows(height, n_bands, k) for k in range(n_bands)]


def gather_strips(dist, strip, full, height, n_bands):
    """All-gather the band strips (rows of `full` owned by each rank) into `full` on every rank.
    `strip` is this rank's rows (a contiguous [rows, W] tensor); `full` is [H, W]."""
    rows = all_bands(height, n_bands)
    if len({b - a for a, b in rows}) == 1:
        dist.all_gather_into_tensor(full, strip)
        return full
    # ragged bands (H not divisible by N): gather strips padded to the tallest band
    import torch

    tallest = max(b - a for a, b in rows)
    padded = torch.zeros((tallest, full.shape[1]), dtype=full.dtype, device=full.device)
    padded[: strip.shape[0]] = strip
    tmp = torch.empty((n_bands * tallest, full.shape[1]), dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(tmp, padded)
    for k, (a, b) in enumerate(rows):
        full[a:b] = tmp[k * tallest : k * tallest + (b - a)]
    return full
