"""N > 1 host logic on CPU: world_size-2 (and 3, ragged) gloo process groups shard a frame by scanline bands,
each rank renders only its band (with the oracle here — no GPU in this test), the strips are all-gathered
with the same helper bench.py uses, and the result must equal the single-band frame bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 256, 190  # 190 rows: ragged for 3 bands


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    from coherence_renderer_b200 import bands, scene, torch_plumbing
    from oracle import pyoracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    objs, n, nbg, edges, points = scene.lion_scene(W, H, 0.55).arrays()
    y0, y1 = bands.band_rows(H, world, rank)
    # bbox_reject=False: the reference's trivial reject compares `bounds_of_basicshape` (which can be one
    # pixel short of the shape) with the bounding box of the update, so the reference ITSELF renders a band
    # differently from the same rows of a whole frame (render.ml:1270-1279; DESIGN.md "known divergences").
    strip = pyoracle.render_frame(objs, n - nbg, nbg, edges, points, (0, y0, W, y1 - y0), bbox_reject=False)
    full = torch.zeros((H, W), dtype=torch.int32)
    torch_plumbing.gather_strips(dist, torch.from_numpy(strip.view(np.int32)).contiguous(), full, H, world)
    if rank == 0:
        np.save(out_path, full.numpy().view(np.uint32))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_band_sharded_frame_equals_whole_frame(tmp_path, world, oracle):
    from coherence_renderer_b200 import bands, scene

    assert bands.all_bands(H, world)[0][0] == 0 and bands.all_bands(H, world)[-1][1] == H
    assert all(a[1] == b[0] for a, b in zip(bands.all_bands(H, world), bands.all_bands(H, world)[1:]))
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    objs, n, nbg, edges, points = scene.lion_scene(W, H, 0.55).arrays()
    whole = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    assert np.array_equal(np.load(out), whole)


def test_balanced_bands_partition_rows_on_cell_boundaries():
    """Cost-weighted bands: contiguous, cover every row once, cut on multiples of the cell
    height, and carry about equal estimated cost."""
    import numpy as np

    from coherence_renderer_b200 import bands, scene

    H, W = 2160, 3840
    _, _, _, edges, _ = scene.lion_scene(W, H, 7.0).arrays()
    cost = bands.row_costs(edges, H, W)
    assert cost.min() > 0 and cost.max() > 5 * cost.min()
    for n in (1, 2, 3, 4, 8):
        bl = bands.balanced_bands(cost, n)
        assert bl[0][0] == 0 and bl[-1][1] == H and all(bl[k][1] == bl[k + 1][0] for k in range(n - 1))
        assert all(b > a for a, b in bl) and all(a % 16 == 0 for a, _ in bl)
        share = [cost[a:b].sum() / cost.sum() for a, b in bl]
        assert max(share) < 1.25 / n
    flat = np.ones(100)
    assert bands.balanced_bands(flat, 3, align=1) == [(0, 33), (33, 67), (67, 100)] or sum(b - a for a, b in bands.balanced_bands(flat, 3, align=1)) == 100
