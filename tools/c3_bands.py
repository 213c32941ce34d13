"""C3 (10^5 objects at 7680x4320) split into scanline bands over the GPUs of one box: torchrun --nproc-per-node N
tools/c3_bands.py.  Every rank renders its band into the symmetric-memory framebuffers of all ranks (fused gather);
prints the per-frame time (max over ranks, CUDA events) on rank 0."""
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
symm = None
if world > 1:
    import torch.distributed._symmetric_memory as symm_mem

    fb = symm_mem.empty((H, W), dtype=torch.int32, device=torch.device("cuda", local))
    fb.zero_()
    symm = symm_mem.rendezvous(fb, dist.group.WORLD)
    ptrs = [int(symm.buffer_ptrs[r]) for r in range(world)]
    ctx.fb_attach(ptrs[rank])
    ctx.fb_set_peers([ptrs[r] for r in range(world) if r != rank])
else:
    fb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    ctx.fb_attach(fb.data_ptr())
sc = ctx.scene_create(objs, nbg, e, p)


def frame():
    ctx.render_frame(sc, (0, 0, W, H))
    if symm is not None:
        symm.barrier()


for _ in range(3):
    frame()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 20
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
for a, b in ev:
    a.record(stream)
    frame()
    b.record(stream)
torch.cuda.synchronize()
t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / K], dtype=torch.float64, device="cuda")
chk = fb[::7, ::5].to(torch.int64).sum().reshape(1)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert lo.item() == hi.item(), "ranks hold different frames"
if rank == 0:
    print(json.dumps({"config": "C3 bands", "n_gpus": world, "ms_per_frame": t.item(), "Mpx_per_s": W * H / t.item() / 1e3, "checksum": int(chk.item())}))
ctx.scene_free(sc)
ctx.close()
if world > 1:
    dist.destroy_process_group()
