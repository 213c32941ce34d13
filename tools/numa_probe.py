"""Does the host side of the read-back sit on the GPU's NUMA node?  D2H bandwidth of a 33 MB pinned buffer allocated
(first-touched) under the default CPU affinity and under NVML's ideal affinity for the device."""
import os
import time

import torch
import pynvml

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
n_cpu = os.cpu_count()
words = (n_cpu + 63) // 64
mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
ideal = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
print("cpus", n_cpu, "default affinity", len(os.sched_getaffinity(0)), "ideal for GPU0", len(ideal), sorted(ideal)[:4], "...")
try:
    print(open("/sys/devices/system/node/online").read().strip(), "numa nodes online")
except OSError:
    pass
dev = torch.empty(3840 * 2160, dtype=torch.int32, device="cuda")


def bw(tag):
    host = torch.empty(3840 * 2160, dtype=torch.int32).pin_memory()
    host.fill_(1)
    for _ in range(5):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(50):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    s = (time.perf_counter() - t) / 50
    print(f"{tag}: {s * 1e3:.3f} ms per 33 MB = {host.numel() * 4 / s / 1e9:.1f} GB/s")


allc = os.sched_getaffinity(0)
bw("default affinity")
if ideal:
    os.sched_setaffinity(0, ideal & allc or allc)
    bw("GPU-local affinity")
    other = allc - ideal
    if other:
        os.sched_setaffinity(0, other)
        bw("remote affinity")
