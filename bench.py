#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 raster hot path (BASELINE.json configs[1], "C2"):
the lion scene at 3840x2160 with correlated-matte antialiasing, cold cache.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, C ABI)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A step is ONE FRAME: scene resident in HBM -> K1 binning + the fused front-to-back walker
(scan conversion, hidden-surface pruning, AA, compositing) -> RGBA8 framebuffer resident in
HBM (+ for N > 1 the NCCL all-gather of the band strips).  One JSON line on stdout.

N > 1: one process per GPU (torchrun), frames shard by horizontal scanline bands, NCCL only
gathers the strips; total work is fixed -> "strong" scaling.
torch is plumbing here (device buffers for NCCL, events, process group), not the product.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SCALE = 3840, 2160, 7.0
WORKLOAD = "C2: lion.pdf scene (132 AA polygons in a Group over a lightgrey background) at 3840x2160, scale 7.0, cold cache"


# Per-kernel ncu summary of one C2 frame of THIS build (ncu --set full --clock-control none; committed with the
# code): dram__bytes_read.sum + dram__bytes_write.sum give roofline.traffic, smsp__inst_executed.sum the issue-slot
# bound.  The 33 MB frame stays in the 126 MB L2 and the scene is L2-resident, so DRAM traffic is below the
# algorithmic bytes.
NCU_SUMMARY = os.path.join("profiles", "r2_ncu_frame_lion4k.csv")


def ncu_summary():
    """(traffic bytes per frame, warp-instructions per frame) summed over the kernels of the frame, or (None, None)."""
    import csv

    path = os.path.join(ROOT, NCU_SUMMARY)
    if not os.path.exists(path):
        return None, None
    traffic, inst = 0.0, 0.0
    for row in csv.reader(open(path)):
        if not row or row[0].startswith("#") or row[0] == "metric":
            continue
        vals = [float(v) for v in row[2:] if v not in ("", "-")]
        if row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            traffic += sum(vals) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[row[1]]
        elif row[0] == "smsp__inst_executed.sum":
            inst += sum(vals)
    return (int(traffic) if traffic else None), (int(inst) if inst else None)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def build_scene():
    from coherence_renderer_b200 import scene

    b = scene.lion_scene(WIDTH, HEIGHT, SCALE)
    return b.arrays()


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, band-parallel over the host cores (the same
# sharding the GPU arm uses; the reference itself is single-threaded OCaml).
# ---------------------------------------------------------------------------------------------
def _cpu_band(args):
    y0, y1 = args
    from oracle import pyoracle

    objs, n, nbg, edges, points = build_scene()
    t = time.perf_counter()
    img = pyoracle.render_frame(objs, n - nbg, nbg, edges, points, (0, y0, WIDTH, y1 - y0))
    return time.perf_counter() - t, int(img[::97, ::89].astype("uint64").sum())


def cpu_frames(n_frames, cores):
    """Render n_frames full C2 frames with the oracle, each split into `cores` bands run in
    parallel processes.  Returns seconds per frame (wall clock)."""
    import multiprocessing as mp

    bands = [(k * HEIGHT // cores, (k + 1) * HEIGHT // cores) for k in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_band, [(0, 8)] * cores)  # warm the workers (imports, tables)
        t = time.perf_counter()
        for _ in range(n_frames):
            pool.map(_cpu_band, bands)
        return (time.perf_counter() - t) / n_frames


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.warmup > 0:
        cpu_frames(args.warmup, cores)
    steps = max(1, args.steps)  # each step is one full C2 frame (about a quarter of a second on 16 cores)
    sec = cpu_frames(steps, cores)
    mpx = WIDTH * HEIGHT / sec / 1e6
    line = {
        "impl": "reference", "metric": "Mpixels/s (complete antialiased frames, scene -> RGBA8 framebuffer)", "value": mpx, "unit": "Mpx/s",
        "frames_per_s": 1.0 / sec, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32 (+f64 crossings)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT},
        "cpu_baseline": {"value": mpx, "unit": "Mpx/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} full C2 frame(s), each split into {cores} horizontal bands rendered by parallel processes of the oracle (C++ restatement of the single-threaded OCaml reference; OCaml is not installable here)"},
        "e2e": {"value": mpx, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    from coherence_renderer_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = world
    objs, n, nbg, edges, points = build_scene()
    n_edges, n_objs = len(edges), n
    from coherence_renderer_b200 import bands

    # equal scanline bands; COH_BALANCED_BANDS=1 cuts them by estimated cost instead (measured slower on the lion,
    # whose rows are evenly loaded: 0.242 vs 0.223 ms at 4 GPUs — the ragged strip exchange costs more than it saves)
    band_list = bands.balanced_bands(bands.row_costs(edges, HEIGHT, WIDTH), N) if N > 1 and os.environ.get("COH_BALANCED_BANDS") else bands.all_bands(HEIGHT, N)
    y0, y1 = band_list[rank]

    ctx = abi.Context(local)
    ctx_sms = torch.cuda.get_device_properties(local).multi_processor_count
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.fb_configure(WIDTH, HEIGHT, y0, y1)
    # N > 1: the band gather.  Preferred: every rank's framebuffer is allocated by the LIBRARY for export
    # (coh_fb_alloc_shared), the 64-byte CUDA IPC handles are exchanged over the process group, every rank maps the
    # others' framebuffers (coh_fb_open_peer -> coh_fb_set_peers) and the rendering kernels store every finished pixel
    # to all of them over NVLink as they go — compute and "collective" are one kernel, followed only by a cross-rank
    # barrier.  If the mapping cannot be set up on this box, fall back to an NCCL all-gather of the strips.
    fb, full, fused, gather = None, None, False, "none (single GPU)"
    signals, epoch = None, [0]
    display_only = os.environ.get("COH_GATHER", "display") != "all"
    if N > 1 and not os.environ.get("COH_NCCL_GATHER"):
        try:
            from coherence_renderer_b200 import torch_plumbing

            handles = torch_plumbing.exchange_ipc_handles(dist, ctx.fb_alloc_shared())
            if display_only:
                # strips go to the display rank.  Rank r > 0 mirrors its band into rank 0's framebuffer
                # only (H/N rows over NVLink per rank instead of (N-1) H/N); rank 0 ends up with the whole frame.
                fb0 = ctx.fb_open_peer(handles[0]) if rank else None
                ctx.fb_set_peers([fb0] if rank else [])
                if not os.environ.get("COH_NCCL_BARRIER"):
                    # frame signals through the shared framebuffers (coh_frame_signal / coh_frame_wait) instead of a
                    # collective: rank r sets counter r of rank 0's framebuffer behind its frame ("band landed"), rank 0
                    # waits for counters 1 .. N-1 and then sets counter 0 of every other framebuffer ("frame consumed"),
                    # which rank r awaits before it stores the next frame
                    signals = {"fb0": fb0, "others": [ctx.fb_open_peer(handles[r]) for r in range(1, N)] if rank == 0 else []}
                gather = "fused into the rendering kernels: every rank stores its band into the display rank's (rank 0) framebuffer over NVLink (CUDA IPC mapping made by the library); frame signals through counters in the shared framebuffers (coh_frame_signal / coh_frame_wait; COH_NCCL_BARRIER=1: an NCCL all-reduce instead); COH_GATHER=all mirrors into every rank"
            else:
                ctx.fb_set_peers([ctx.fb_open_peer(handles[r]) for r in range(N) if r != rank])
                gather = "fused into the rendering kernels: peer stores over NVLink into every rank's framebuffer (CUDA IPC mappings made by the library) + cross-rank barrier"
            fused = True
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print(f"bench.py: peer framebuffers unavailable ({type(exc).__name__}: {exc}); using NCCL all-gather", file=sys.stderr)
            fused = False
    if not fused and N > 1:
        from coherence_renderer_b200 import torch_plumbing

        fb = torch.zeros((HEIGHT, WIDTH), dtype=torch.int32, device="cuda")  # torch-owned so NCCL can gather it
        ctx.fb_attach(fb.data_ptr())
        full = torch.zeros((HEIGHT, WIDTH), dtype=torch.int32, device="cuda")
        gather = "NCCL all-gather of the band strips after the frame"
    scene_h = ctx.scene_create(objs, nbg, edges, points)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    update = (0, 0, WIDTH, HEIGHT)

    sync_t = torch.zeros(1, dtype=torch.int32, device="cuda")

    def frame_done():
        """Behind the frame's kernels: the cross-rank step that makes the display rank's framebuffer complete."""
        if signals is not None:
            e = epoch[0]
            if rank:
                ctx.frame_signal([signals["fb0"]], rank, e)            # my band has landed in rank 0's framebuffer
            else:
                ctx.frame_wait(list(range(1, N)), e)                   # every band has landed
                ctx.frame_signal(signals["others"], 0, e)              # the frame is complete: the next one may be stored
        elif fused:
            dist.all_reduce(sync_t)  # cross-rank barrier on the stream: every rank's band has landed in this rank's framebuffer
        elif N > 1:
            torch_plumbing.gather_strips(dist, fb[y0:y1], full, HEIGHT, N, rows=band_list)

    def frame_begin():
        if signals is not None:
            epoch[0] += 1
            if rank and epoch[0] > 1:
                ctx.frame_wait([0], epoch[0] - 1)                      # rank 0 is done with the frame before

    def frame():
        frame_begin()
        ctx.render_frame(scene_h, update)
        frame_done()

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()
    ctx.sync()  # surfaces kernel-side errors
    if N > 1 and fused and display_only:
        # the display rank must hold every band exactly as the rank that rendered it holds it
        mine = torch.from_numpy(ctx.fb_read_rgba(0, y0, WIDTH, y1 - y0).view(np.int32)).cuda()
        sums = torch.zeros(N, dtype=torch.int64, device="cuda")
        sums[rank] = mine[::5, ::7].to(torch.int64).sum()
        dist.all_reduce(sums)
        bad = torch.zeros(1, dtype=torch.int32, device="cuda")
        if rank == 0:
            whole = torch.from_numpy(ctx.fb_read_rgba(0, 0, WIDTH, HEIGHT).view(np.int32)).cuda()
            for k, (a, b) in enumerate(band_list):
                if whole[a:b][::5, ::7].to(torch.int64).sum().item() != sums[k].item() or (whole[a:b][:: max((b - a) // 4, 1), 5] == 0).any().item():
                    bad += 1
            del whole
        dist.all_reduce(bad)
        if bad.item():
            raise SystemExit("bench.py: the display rank's frame differs from the bands the ranks rendered")
    elif N > 1:  # every rank must now hold the same, complete frame
        whole = torch.from_numpy(ctx.fb_read_rgba(0, 0, WIDTH, HEIGHT).view(np.int32)).cuda() if fused else full
        chk = whole[::13, ::7].to(torch.int64).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if lo.item() != hi.item() or (whole[:: HEIGHT // 8, 5] == 0).any().item():
            raise SystemExit("bench.py: the gathered frames differ between ranks or have missing bands")
        del whole

    # ---- device-timed steps: L2 flushed (untimed) before every step, CUDA events on the launching stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.set_timing(True)
    l0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev_mid = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]   # after this rank's own kernels, before the barrier
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record(stream)
        frame_begin()
        ctx.render_frame(scene_h, update)
        ev_mid[s].record(stream)
        frame_done()
        ev[s][1].record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    walk_ms, bin_ms, timed = ctx.get_timing()
    ctx.set_timing(False)
    total_ms = sum(step_ms)
    own_ms = sum(a.elapsed_time(m) for (a, _), m in zip(ev, ev_mid)) / args.steps        # this rank's kernels
    wait_ms = sum(m.elapsed_time(b) for (_, b), m in zip(ev, ev_mid)) / args.steps       # barrier / gather: waiting for the slowest rank
    t = torch.tensor([total_ms, walk_ms], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros((N, 4), dtype=torch.float64, device="cuda")
    per_rank[rank] = torch.tensor([bin_ms, walk_ms, own_ms, wait_ms], dtype=torch.float64)
    if N > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank)
    total_ms, walk_ms_max = t.tolist()
    per_rank = per_rank.tolist()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public C-ABI call with HOST buffers: per step the scene is uploaded
    # from host arrays (H2D), rendered, and the band strip read back to pinned host memory (D2H)
    # (frame k's read-back overlaps frame k+1's upload and rendering: two pinned host buffers, the
    # library's asynchronous read; every copy has completed before the clock stops)
    # N > 1: the WHOLE frame ends up in one pinned host buffer on rank 0 (every framebuffer holds every band after the
    # barrier); the other ranks upload their scene and render, and read nothing back
    reader = rank == 0
    ry0, ry1 = (0, HEIGHT) if (N > 1 and fused) else (y0, y1)
    hosts = [torch.empty((ry1 - ry0, WIDTH), dtype=torch.int32).pin_memory() for _ in range(2)] if (reader or not fused) else []
    hosts_np = [h.numpy().view(np.uint32) for h in hosts]
    e2e_steps = max(3, min(args.steps, 200))   # (the first frame's render is not hidden behind a read-back: pipeline fill, amortised)
    h2d = objs._length_ * abi.C.sizeof(abi.CohObject) + edges.nbytes + points.nbytes
    d2h = hosts_np[0].nbytes if hosts_np else 0

    def e2e_run(k):
        for i in range(k):
            sh = ctx.scene_create(objs, nbg, edges, points)
            ctx.render_frame(sh, update)
            if fused:
                dist.all_reduce(sync_t)   # (stream-ordered: the read-back below is queued behind it)
            if hosts_np:
                ctx.fb_read_rgba_async(0, ry0, WIDTH, ry1 - ry0, hosts_np[i & 1])
            ctx.scene_free(sh)
        ctx.fb_read_wait()

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = t.item()

    if rank == 0:
        ms = total_ms / args.steps
        mpx = WIDTH * HEIGHT / (ms * 1e-3) / 1e6
        peak, peak_src = peaks()
        traffic, inst = ncu_summary()
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6 if clocks else 1965.0e6
        # algorithmic bytes of ONE walker launch on this rank: one RGBA8 write per output
        # pixel of the band + one read of every prepared edge (32 B) and object record
        alg_bytes = 4 * WIDTH * (y1 - y0) + 16 * n_edges + 32 * n_objs
        ach = alg_bytes / (walk_ms_max * 1e-3) / 1e9 if walk_ms_max > 0 else 0.0
        line = {
            "metric": "Mpixels/s (complete antialiased frames, scene -> RGBA8 framebuffer)", "value": mpx, "unit": "Mpx/s",
            "frames_per_s": 1e3 / ms, "n_gpus": N, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32 (+f64 crossings)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "bands": N, "band_rows": [list(b) for b in band_list], "l2": "256 MB flush write before every timed step (untimed)",
                       "step": "one frame = 5 launches: background prefill (second stream) beside scan conversion, visibility and antialiasing of every (cell entry, row) pair, then the row compositor; the cell lists are kept with the scene (binned once)", "gather": gather},
            "roofline": {"bound": "hbm", "kernel": "raster phase: k_pre_scan + k_pre_vis + k_pre_aa_runs (dominant) + k_comp_rows", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic if N == 1 else None, "traffic_source": NCU_SUMMARY if traffic else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": walk_ms_max, "binning_ms": bin_ms,
                         # the honest second bound: the path is integer / bit work, limited by instruction issue, not by HBM
                         "issue_frac": (inst / (ctx_sms * 4 * sm_hz * ms * 1e-3)) if (inst and N == 1) else None,
                         "warp_instructions_per_frame": inst if N == 1 else None,
                         "note": "kernel_ms is the device time of the raster-phase launches together (CUDA events on the launching stream); integer/bit + FP64-crossing work, issue bound, not HBM bound (DESIGN.md): issue_frac = warp-instructions of the frame (ncu, same build) / (SMs x 4 schedulers x SM clock x step time)"},
            "e2e": {"value": WIDTH * HEIGHT / e2e_s / 1e6, "unit": "Mpx/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "path": "per frame: coh_scene_create(host arrays) + coh_render_frame + coh_fb_read_rgba_async(pinned host); the read-back of frame k overlaps frame k+1, all copies complete inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if N > 1:
            # where a step goes on every rank (ms, averaged over the timed steps): K1 binning and raster phase from the
            # library's own CUDA events, this rank's kernels end to end, and the wait at the cross-rank barrier
            line["per_rank"] = [{"rank": r, "band_rows": list(band_list[r]), "binning_ms": v[0], "raster_ms": v[1], "own_kernels_ms": v[2], "barrier_wait_ms": v[3]} for r, v in enumerate(per_rank)]
            line["e2e"]["path"] = ("per frame: every rank uploads the scene (coh_scene_create) and renders its band; after the cross-rank barrier rank 0 reads the WHOLE frame "
                                   "(every framebuffer holds every band) into pinned host memory with coh_fb_read_rgba_async; the read-back of frame k overlaps frame k+1")
        if args.cpu_baseline:
            cores = os.cpu_count() or 1
            sec = cpu_frames(2, cores)
            sec1 = cpu_frames(1, 1)   # the reference is single-threaded: one frame on one core
            line["cpu_baseline"] = {"value": WIDTH * HEIGHT / sec / 1e6, "unit": "Mpx/s", "cores": cores, "kind": "port", "ms_per_frame": sec * 1e3,
                                    "cores_1_ms_per_frame": sec1 * 1e3, "cores_1_value": WIDTH * HEIGHT / sec1 / 1e6,
                                    "sample": f"2 full C2 frames, each split into {cores} horizontal bands rendered in parallel by the oracle (C++ restatement of the single-threaded OCaml reference), and 1 full frame on 1 core"}
        print(json.dumps(line))
    ctx.scene_free(scene_h)
    ctx.close()
    if N > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus > 1:
            args.cpu_baseline = False  # rank 0 at N=1 only
        run_ours(args)


if __name__ == "__main__":
    main()
