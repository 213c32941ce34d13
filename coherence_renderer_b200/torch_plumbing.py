"""Optional plumbing for hosts that run one process per GPU under torch.distributed (bench.py under torchrun, the
gloo tests): exchange of the framebuffers' CUDA IPC handles, and an NCCL / gloo all-gather of band strips for boxes
where peer mapping is not available.  Nothing in the library or in the rest of this package needs torch."""
from .bands import all_bands


def exchange_ipc_handles(dist, handle):
    """All-gather the 64-byte CUDA IPC handle of every rank's shared framebuffer (host-side, any backend)."""
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(handle))
    return out


def gather_strips(dist, strip, full, height, n_bands, rows=None):
    """All-gather the band strips (rows of `full` owned by each rank) into `full` on every rank.
    `strip` is this rank's rows (a contiguous [rows, W] tensor); `full` is [H, W]; `rows` = the bands
    (default: the equal split)."""
    explicit = rows is not None
    rows = rows or all_bands(height, n_bands)
    if len({b - a for a, b in rows}) == 1:
        dist.all_gather_into_tensor(full, strip)
        return full
    if explicit and dist.get_backend() == "nccl":
        # cost-balanced bands have very different heights: one grouped exchange of exact-size strips
        # (ncclGroupStart / Send / Recv / End) instead of padding every strip to the tallest band
        rank = dist.get_rank()
        ops = []
        for k, (a, b) in enumerate(rows):
            if k == rank:
                continue
            ops.append(dist.P2POp(dist.isend, strip, k))
            ops.append(dist.P2POp(dist.irecv, full[a:b], k))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        a, b = rows[rank]
        if full[a:b].data_ptr() != strip.data_ptr():
            full[a:b] = strip
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
