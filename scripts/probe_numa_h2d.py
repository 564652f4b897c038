#!/usr/bin/env python
"""Which host cores should first-touch a rank's page-locked staging buffers?  The GPU boxes hide
their NUMA topology (numa_node = -1), so measure it: for every visible GPU and every quarter of
the allowed CPUs, pin a buffer while running on those CPUs and time the host->device copy, alone
and with all GPUs copying at once."""
import os, sys, time, threading
import torch

n_gpu = torch.cuda.device_count()
allowed = sorted(os.sched_getaffinity(0))
print("gpus", n_gpu, "cpus", len(allowed), allowed[:4], "...", allowed[-4:], flush=True)
try:
    print("nodes:", sorted(p for p in os.listdir("/sys/devices/system/node") if p.startswith("node")))
except OSError as e:
    print("no /sys/devices/system/node:", e)
os.system("lscpu | grep -i -E 'socket|numa|model name|^CPU\\(s\\)' ; free -g | head -2; nvidia-smi topo -m 2>/dev/null | head -16")
NB = 512 << 20
parts = 4 if len(allowed) >= 8 else 2
sets = [allowed[i * len(allowed) // parts:(i + 1) * len(allowed) // parts] for i in range(parts)]
bufs = {}
for si, cs in enumerate(sets):
    os.sched_setaffinity(0, cs)
    t = torch.empty(NB, dtype=torch.uint8).pin_memory()
    t.fill_(1)
    bufs[si] = t
os.sched_setaffinity(0, allowed)
dev = [torch.empty(NB, dtype=torch.uint8, device=f"cuda:{g}") for g in range(n_gpu)]


def bw(g, src, reps=4):
    s = torch.cuda.Stream(device=g)
    with torch.cuda.stream(s):
        dev[g].copy_(src, non_blocking=True)
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dev[g].copy_(src, non_blocking=True)
        s.synchronize()
    return NB * reps / (time.perf_counter() - t0) / 1e9


best = {}
for g in range(n_gpu):
    row = [bw(g, bufs[si]) for si in range(parts)]
    best[g] = max(range(parts), key=lambda i: row[i])
    print(f"gpu{g} alone: " + "  ".join(f"cpuset{si}={v:.1f}" for si, v in enumerate(row)) + f"  best={best[g]}", flush=True)


def concurrent(choice, label):
    out = [0.0] * n_gpu
    # each GPU needs its own source buffer: allocate under the chosen cpuset
    srcs = []
    for g in range(n_gpu):
        os.sched_setaffinity(0, sets[choice[g]])
        t = torch.empty(NB, dtype=torch.uint8).pin_memory(); t.fill_(g); srcs.append(t)
    os.sched_setaffinity(0, allowed)
    th = [threading.Thread(target=lambda g=g: out.__setitem__(g, bw(g, srcs[g], reps=8))) for g in range(n_gpu)]
    [t.start() for t in th]; [t.join() for t in th]
    print(f"{label}: total {sum(out):.1f} GB/s  per gpu " + " ".join(f"{v:.1f}" for v in out), flush=True)


concurrent({g: 0 for g in range(n_gpu)}, "all from cpuset0")
concurrent({g: parts - 1 for g in range(n_gpu)}, f"all from cpuset{parts - 1}")
concurrent({g: g * parts // n_gpu for g in range(n_gpu)}, "spread by gpu index")
concurrent(best, "best-alone per gpu")
