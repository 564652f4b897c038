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
