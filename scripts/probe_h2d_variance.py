#!/usr/bin/env python
"""Why does the page-locked e2e of one GPU read 55.5 GB/s in most processes and 46-48 GB/s in some
(gpurun_out/r02x_bench.json vs r02x_bench_c128.json, same box, same GPU)?  Placement of the buffer
(NUMA node first touched) or other traffic on the host fabric?  Prints the topology the container
sees, where the pages of page-locked buffers made under different CPU sets live (move_pages), and
the host->device rate of each buffer over ~1.5 s windows, interleaved, several rounds."""
import ctypes, os, time
import torch

allowed = sorted(os.sched_getaffinity(0))
print("cpus allowed", len(allowed), allowed, flush=True)
os.system("grep -E 'Cpus_allowed_list|Mems_allowed_list' /proc/self/status; ls /sys/devices/system/node 2>&1 | tr '\\n' ' '; echo; "
          "for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemFree $n/meminfo); done; "
          "lscpu | grep -i -E 'socket|numa|model name|^CPU\\(s\\)|thread'; nvidia-smi topo -m 2>/dev/null | head -6; "
          "cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c")
libc = ctypes.CDLL(None, use_errno=True)


def nodes_of(t, samples=64):
    """NUMA node of `samples` pages of tensor t (move_pages with nodes=NULL queries)."""
    n = t.numel()
    addrs = (ctypes.c_void_p * samples)(*[t.data_ptr() + (i * (n // samples) // 4096) * 4096 for i in range(samples)])
    status = (ctypes.c_int * samples)()
    rc = libc.syscall(279, 0, samples, addrs, None, status, 0)
    if rc != 0:
        return "move_pages errno %d" % ctypes.get_errno()
    out = {}
    for s in status:
        out[s] = out.get(s, 0) + 1
    return out


NB = 1 << 30
half = len(allowed) // 2
sets = {"all": allowed, "low": allowed[:half], "high": allowed[half:]}
bufs = {}
for name, cs in sets.items():
    os.sched_setaffinity(0, cs)
    t = torch.empty(NB, dtype=torch.uint8).pin_memory()
    t.fill_(1)
    bufs[name] = t
    print("buffer", name, "made on cpu", os.sched_getcpu() if hasattr(os, "sched_getcpu") else "?", "pages on nodes", nodes_of(t), flush=True)
os.sched_setaffinity(0, allowed)
dev = torch.empty(NB, dtype=torch.uint8, device="cuda:0")
s = torch.cuda.Stream()


def bw(src, seconds=1.5):
    with torch.cuda.stream(s):
        dev.copy_(src, non_blocking=True)
        s.synchronize()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(8):
                dev.copy_(src, non_blocking=True)
            s.synchronize()
            n += 8
    return NB * n / (time.perf_counter() - t0) / 1e9


for rnd in range(4):
    print("round", rnd, "  ".join(f"{name}={bw(t):.1f}" for name, t in bufs.items()), "GB/s", flush=True)
# the library's allocator (cudaMallocHost) instead of torch's
from bandup_b200.unfold import pinned_empty, pinned_free  # noqa: E402
import numpy as np  # noqa: E402
for name, cs in sets.items():
    os.sched_setaffinity(0, cs)
    h = pinned_empty(NB)
    np.asarray(h).view(np.uint8)[:] = 1
    t = torch.from_numpy(np.asarray(h).view(np.uint8))
    print("cudaMallocHost under", name, "pages on nodes", nodes_of(t), f"{bw(t):.1f} GB/s", flush=True)
    pinned_free(h)
os.sched_setaffinity(0, allowed)
