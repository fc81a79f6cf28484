"""Host-side staging helpers for the end-to-end path: NUMA-local pinned buffers.

Pinned host memory is physically placed by first touch.  On an 8-GPU box every rank's staging buffers otherwise
end up on whichever node the launcher started the process on, and all H2D / D2H traffic of the box funnels
through one memory controller (round-1 finding: end-to-end scaling 0.67 at 8 GPUs with every rank on NUMA 0).
`bind_to_gpu_numa` pins the calling process to the CPUs that are local to its GPU's PCIe root *before* the
buffers are allocated, so that first touch puts them next to the GPU.
"""
from __future__ import annotations

import os
import subprocess
from typing import Optional


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_local_cpus(index: int) -> Optional[set]:
    """CPUs local to GPU `index` (sysfs `local_cpulist` of its PCI function), or None when unknown."""
    try:
        out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not out:
            return None
        dom, rest = out.split(":", 1)
        bdf = f"{dom[-4:]}:{rest}"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        return None


def gpu_numa_node(index: int) -> Optional[int]:
    try:
        out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        dom, rest = out.split(":", 1)
        with open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node") as f:
            return int(f.read().strip())
    except Exception:
        return None


_orig_affinity: Optional[set] = None


def restore_affinity() -> None:
    """Undo bind_to_gpu_numa (host-side legs that should see every core, e.g. a CPU baseline)."""
    global _orig_affinity
    if _orig_affinity is not None:
        try:
            os.sched_setaffinity(0, _orig_affinity)
        except OSError:
            pass
        _orig_affinity = None


def bind_to_gpu_numa(index: int) -> dict:
    """Restrict this process to the CPUs local to GPU `index` (no-op when the topology is unknown or the local
    set does not intersect the CPUs we are allowed to use).  Returns what was done, for the bench record."""
    info = {"gpu": index, "numa_node": gpu_numa_node(index), "bound": False}
    cpus = gpu_local_cpus(index)
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return info
    info["cpus_before"] = len(allowed)
    global _orig_affinity
    if _orig_affinity is None:
        _orig_affinity = set(allowed)
    if cpus:
        target = cpus & allowed
        if target and target != allowed:
            try:
                os.sched_setaffinity(0, target)
                info["bound"] = True
            except OSError:
                pass
        info["cpus_local"] = len(target)
    return info
