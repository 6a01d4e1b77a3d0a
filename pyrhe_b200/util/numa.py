"""Keep a rank's host threads (and therefore its pinned staging memory, which is placed on the node of the thread that
allocates it) on the NUMA node its GPU hangs off.  One process per GPU streams its `.bed` share through pinned host
memory (`RheEngine.stream_genotypes`); with eight ranks on a two-socket box a rank whose ring lies on the far socket
copies across the socket interconnect and the step waits for the slowest rank.  Linux sysfs only; every failure is a
no-op that reports why."""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> int:
    """NUMA node of a CUDA device from sysfs (-1: unknown / single node)."""
    import torch
    p = torch.cuda.get_device_properties(device_index)
    bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
        return int(fh.read().strip())


def bind_to_gpu_node(device_index: int, min_cpus: int = 2) -> dict:
    """Restrict this process to the allowed CPUs of the GPU's NUMA node.  Returns a report (`bound`, `node`, `cpus`,
    `why`); leaves the affinity alone when the node is unknown or the allowed CPUs on it are fewer than `min_cpus`."""
    report = {"bound": False, "node": -1, "cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 0}
    try:
        node = gpu_numa_node(device_index)
        report["node"] = node
        if node < 0:
            report["why"] = "GPU reports no NUMA node"
            return report
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            local = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        mine = allowed & local
        if len(mine) < min_cpus:
            report["why"] = f"{len(mine)} of the {len(allowed)} allowed CPUs are on node {node}"
            return report
        if mine != allowed:
            os.sched_setaffinity(0, mine)
        report.update(bound=True, cpus=len(mine))
    except Exception as exc:                                    # sysfs absent, permissions, non-Linux
        report["why"] = f"{type(exc).__name__}: {exc}"
    return report
