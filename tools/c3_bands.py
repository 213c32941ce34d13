"""C3 (10^5 objects at 7680x4320) split into scanline bands over the GPUs of one box: torchrun --nproc-per-node N
tools/c3_bands.py.  Every rank renders its band into the framebuffers of all ranks (the library's CUDA IPC mappings:
coh_fb_alloc_shared / coh_fb_open_peer / coh_fb_set_peers — the gather is fused into the rendering kernels), followed
by a cross-rank barrier on the stream; prints the per-frame time (max over ranks, CUDA events) on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from coherence_renderer_b200 import abi, bands, scene

W, H, NOBJ = 7680, 4320, 100000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
objs, n, nbg, e, p = scene.random_scene(W, H, NOBJ).arrays()
y0, y1 = bands.band_rows(H, world, rank)
ctx = abi.Context(local)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ctx.fb_configure(W, H, y0, y1)
fused = world > 1
sync_t = torch.zeros(1, dtype=torch.int32, device="cuda")
if fused:
    from coherence_renderer_b200 import torch_plumbing

    handles = torch_plumbing.exchange_ipc_handles(dist, ctx.fb_alloc_shared())
    if os.environ.get("COH_GATHER", "display") != "all":   # strips to the display rank (rank 0) only
        ctx.fb_set_peers([ctx.fb_open_peer(handles[0])] if rank else [])
    else:
        ctx.fb_set_peers([ctx.fb_open_peer(handles[r]) for r in range(world) if r != rank])
sc = ctx.scene_create(objs, nbg, e, p)


def frame():
    ctx.render_frame(sc, (0, 0, W, H))
    if fused:
        dist.all_reduce(sync_t)   # every rank's band has landed in this rank's framebuffer


for _ in range(3):
    frame()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 20
ctx.set_timing(True)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
for a, b in ev:
    a.record(stream)
    frame()
    b.record(stream)
torch.cuda.synchronize()
walk_ms, bin_ms, _ = ctx.get_timing()
ctx.set_timing(False)
t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / K], dtype=torch.float64, device="cuda")
per_rank = torch.zeros((world, 2), dtype=torch.float64, device="cuda")
per_rank[rank] = torch.tensor([bin_ms, walk_ms], dtype=torch.float64)
if world > 1:
    dist.all_reduce(per_rank)
import numpy as np  # noqa: E402

display = world > 1 and os.environ.get("COH_GATHER", "display") != "all"
if display:
    sums = torch.zeros(world, dtype=torch.int64, device="cuda")
    sums[rank] = int(ctx.fb_read_rgba(0, y0, W, y1 - y0)[::7, ::5].astype(np.uint64).sum()) & 0x7FFFFFFFFFFFFFFF
    dist.all_reduce(sums)
chk = torch.tensor([int(ctx.fb_read_rgba(0, 0, W, H)[::7, ::5].astype(np.uint64).sum()) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if display:
        if rank == 0:
            whole = ctx.fb_read_rgba(0, 0, W, H)
            for k in range(world):
                a, b = bands.band_rows(H, world, k)
                assert (int(whole[a:b][::7, ::5].astype(np.uint64).sum()) & 0x7FFFFFFFFFFFFFFF) == sums[k].item(), "the display rank's frame differs from band %d" % k
    else:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert lo.item() == hi.item(), "ranks hold different frames"
if rank == 0:
    print(json.dumps({"config": "C3 bands", "n_gpus": world, "ms_per_frame": t.item(), "Mpx_per_s": W * H / t.item() / 1e3, "checksum": int(chk.item()),
                      "per_rank_binning_raster_ms": [[round(v, 4) for v in r] for r in per_rank.tolist()]}))
ctx.scene_free(sc)
ctx.close()
if world > 1:
    dist.destroy_process_group()
