"""Host-side placement for the end-to-end path: bind a rank's threads to the CPUs of the NUMA node its GPU hangs off,
BEFORE pinned staging buffers are allocated, so that host->device copies read node-local memory (first touch).

On an 8-GPU box torchrun starts eight ranks wherever the scheduler puts them; a rank whose pinned buffers live on the
far socket pushes every byte of its H2D stream across the inter-socket link, which is what bounds the aggregate
end-to-end rate (bench.py `e2e` at N = 8).  No external tools: the topology is read from sysfs.
"""
from __future__ import annotations

import os
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


def gpu_local_cpus(device_index: int) -> Optional[List[int]]:
    """CPUs local to the PCIe device of CUDA device `device_index` (sysfs local_cpulist), or None if unknown."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the CPUs local to its GPU (intersection with the CPUs it may already use).
    Returns {'bound': bool, 'cpus': n, 'why': str}; never raises -- an unknown topology leaves the affinity alone."""
    cpus = gpu_local_cpus(device_index)
    if not cpus:
        return {"bound": False, "cpus": len(os.sched_getaffinity(0)), "why": "no local_cpulist for the device"}
    allowed = os.sched_getaffinity(0)
    want = sorted(set(cpus) & allowed)
    if not want:
        return {"bound": False, "cpus": len(allowed), "why": "GPU-local CPUs are outside the allowed set"}
    if set(want) == set(allowed):
        return {"bound": False, "cpus": len(allowed), "why": "single NUMA node (every allowed CPU is GPU-local)"}
    try:
        os.sched_setaffinity(0, want)
    except OSError as e:
        return {"bound": False, "cpus": len(allowed), "why": f"sched_setaffinity failed: {e}"}
    return {"bound": True, "cpus": len(want), "why": "bound to the GPU's NUMA-local CPUs"}
