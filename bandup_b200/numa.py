"""Bind a rank to the CPU cores (and therefore, by first touch, the host memory) of the NUMA node
its GPU hangs off.  End to end the unfolding path is bound by the host->device copy of the
coefficient blocks; page-locked staging buffers that live on the far socket halve that copy."""
from __future__ import annotations

import os
from pathlib import Path
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(pci_bus_id: str) -> Optional[int]:
    """NUMA node of a PCI device ('0000:1b:00.0'), or None when the kernel does not say."""
    p = Path("/sys/bus/pci/devices") / pci_bus_id.lower() / "numa_node"
    try:
        n = int(p.read_text().strip())
        return n if n >= 0 else None
    except (OSError, ValueError):
        return None


def node_cpus(node: int) -> List[int]:
    try:
        return _parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text())
    except OSError:
        return []


def bind_to_gpu_node(device_index: int, ranks_on_node_hint: int = 1, local_rank: int = 0) -> dict:
    """sched_setaffinity to the cores of the GPU's NUMA node (intersected with what the process is
    allowed to use).  Returns what was done, for the bench record.  Never raises."""
    info = {"bound": False}
    try:
        import torch
        prop = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        node = gpu_numa_node(bus)
        info.update(pci=bus, numa_node=node)
        if node is None:
            return info
        allowed = set(os.sched_getaffinity(0))
        cpus = sorted(allowed.intersection(node_cpus(node)))
        if not cpus:
            return info
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, n_cpus=len(cpus))
    except Exception as e:  # pragma: no cover - best effort
        info["error"] = repr(e)
    return info


# ---- which GPU a local rank should use -----------------------------------------------------------
# End to end the path is bound by the host->device copy of the coefficient blocks.  On the 8-GPU
# boxes pairs of GPUs hang off one PCIe switch and share its uplink: four ranks on GPUs 0-3 get
# ~29 GB/s each, four ranks on one GPU per switch ~55 GB/s each.  So local ranks are spread over
# the PCIe tree (farthest first) instead of filling it in index order.

def pci_path(pci_bus_id: str) -> List[str]:
    """Bridges between the root complex and the device, from sysfs ('pci0000:15', '0000:15:01.0', ...);
    empty when sysfs does not show the device."""
    try:
        real = os.path.realpath(str(Path("/sys/bus/pci/devices") / pci_bus_id.lower()))
    except OSError:
        return []
    parts = [p for p in real.split("/") if p.startswith("pci") or p.count(":") == 2]
    return parts if len(parts) >= 2 else []


def _common_prefix(a: List[str], b: List[str]) -> int:
    n = 0
    for x, y in zip(a, b):
        if x != y:
            break
        n += 1
    return n


def spread_order(paths: List[List[str]], numa_nodes: Optional[List[Optional[int]]] = None) -> List[int]:
    """Order of the devices such that every prefix of it is spread over the PCIe tree as widely as
    possible: greedy farthest-first on (same NUMA node, shared bridges); ties go to the lower index.
    Without usable paths: bit-reversed index order (neighbours share a switch on the HGX boards)."""
    n = len(paths)
    if n <= 1:
        return list(range(n))
    if not all(paths):
        bits = max(1, (n - 1).bit_length())
        rev = sorted(range(1 << bits), key=lambda i: int(format(i, f"0{bits}b")[::-1], 2))
        return [i for i in rev if i < n]
    numa_nodes = numa_nodes or [None] * n

    def closeness(i: int, j: int) -> int:       # larger = more shared hardware between the two
        c = _common_prefix(paths[i], paths[j])
        if numa_nodes[i] is not None and numa_nodes[i] == numa_nodes[j]:
            c += 1
        return c
    order = [0]
    while len(order) < n:
        rest = [i for i in range(n) if i not in order]
        # the candidate whose CLOSEST already-chosen device is the least close; then the one that is
        # least close to all of them in total
        best = min(rest, key=lambda i: (max(closeness(i, j) for j in order), sum(closeness(i, j) for j in order), i))
        order.append(best)
    return order


def device_for_local_rank(local_rank: int, local_world: int) -> dict:
    """The CUDA device a local rank should use and why: {'device', 'order', 'pci', 'source'}.
    Deterministic: every rank computes the same order.  BANDUP_B200_GPU_ORDER=0,4,2,... overrides."""
    import torch
    n = torch.cuda.device_count()
    info = {"device": local_rank % max(n, 1), "order": list(range(n)), "source": "index order"}
    try:
        env = os.environ.get("BANDUP_B200_GPU_ORDER")
        bus = []
        for i in range(n):
            prop = torch.cuda.get_device_properties(i)
            bus.append("%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id))
        if env:
            order = [int(x) for x in env.split(",")]
            source = "BANDUP_B200_GPU_ORDER"
        elif local_world >= n:
            order, source = list(range(n)), "all visible GPUs in use: index order"
        else:
            paths = [pci_path(b) for b in bus]
            order = spread_order(paths, [gpu_numa_node(b) for b in bus])
            source = "PCIe tree from sysfs, farthest first" if all(paths) else "no sysfs PCI paths: bit-reversed index order"
        if sorted(order) != list(range(n)):
            raise ValueError(f"GPU order {order} is not a permutation of 0..{n - 1}")
        info.update(device=order[local_rank % n], order=order, pci=bus, source=source)
    except Exception as e:  # pragma: no cover - best effort
        info["error"] = repr(e)
    return info


# ---- measured host-link sharing -----------------------------------------------------------------------
# On virtualised boxes sysfs shows a flat PCI bus, so the tree above says nothing.  The copy engines
# do.  On the 8-GPU boxes here any two GPUs copy host->device at their full 55 GB/s together, but GPUs
# 0-3 share a ~115 GB/s path and GPUs 4-7 another: four ranks on GPUs 0-3 get 29 GB/s each, four
# ranks on {0, 1, 4, 5} the full rate.  probe_h2d_order() finds such an order without knowing the
# topology: starting from GPU 0 it keeps adding the GPU with which the set's TOTAL measured rate
# (all copying at once) is highest, lowest index on ties.  28 short copy tests for 8 GPUs; run it in a
# throw-away process (`python -m bandup_b200.numa --probe`) so that the contexts it opens on every
# GPU die with it.

def greedy_order(n: int, total_rate, tie: float = 0.03):
    """GPU 0, then always the GPU with which the chosen set's total rate (total_rate(list of devices),
    all active at once) is highest; a candidate must beat the best so far by more than `tie` to
    displace a lower index.  Returns (order, total rate of every prefix)."""
    if n <= 0:
        return [], []
    order = [0]
    prefix_total = [float(total_rate([0]))]
    while len(order) < n:
        best_g, best_t = -1, -1.0
        for g in range(n):
            if g in order:
                continue
            t = float(total_rate(order + [g]))
            if best_g < 0 or t > best_t * (1.0 + tie):
                best_g, best_t = g, t
        order.append(best_g)
        prefix_total.append(best_t)
    return order, prefix_total


def probe_h2d_order(mbytes: int = 96, tie: float = 0.03) -> dict:
    import torch
    n = torch.cuda.device_count()
    nbytes = int(mbytes) << 20
    host, devb, streams = [], [], []
    for i in range(n):
        with torch.cuda.device(i):
            host.append(torch.empty(nbytes, dtype=torch.uint8, pin_memory=True))
            devb.append(torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}"))
            streams.append(torch.cuda.Stream(device=i))

    def rates(devs, reps=2):
        """GB/s of every device in `devs`, all copying at once (best of `reps`)."""
        best = [0.0] * len(devs)
        for _ in range(reps):
            ev = []
            for d in devs:
                torch.cuda.synchronize(d)
            for d in devs:
                with torch.cuda.device(d), torch.cuda.stream(streams[d]):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    devb[d].copy_(host[d], non_blocking=True)
                    devb[d].copy_(host[d], non_blocking=True)
                    b.record()
                    ev.append((a, b))
            for k, d in enumerate(devs):
                torch.cuda.synchronize(d)
                best[k] = max(best[k], 2 * nbytes / (ev[k][0].elapsed_time(ev[k][1]) * 1e-3) / 1e9)
        return best

    for i in range(n):
        rates([i], reps=1)                      # warm the contexts and the pinned mappings
    solo = [rates([i])[0] for i in range(n)]
    order, prefix_total = greedy_order(n, lambda devs: sum(rates(devs)), tie)
    return {"solo_GBps": [round(x, 1) for x in solo], "order": order,
            "prefix_total_GBps": [round(x, 1) for x in prefix_total],
            "source": "measured: greedy order by total host->device rate of the GPUs copying at once "
                      "(bandup_b200.numa.probe_h2d_order)"}


def probed_gpu_order(timeout_s: float = 120.0) -> Optional[dict]:
    """probe_h2d_order() in a child process; None when it fails."""
    import json
    import subprocess
    import sys
    try:
        r = subprocess.run([sys.executable, "-m", "bandup_b200.numa", "--probe"], capture_output=True, text=True,
                           timeout=timeout_s, cwd=str(Path(__file__).resolve().parent.parent))
        for line in reversed(r.stdout.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        pass
    return None


if __name__ == "__main__":
    import json
    import sys
    if "--probe" in sys.argv:
        print(json.dumps(probe_h2d_order()), flush=True)
