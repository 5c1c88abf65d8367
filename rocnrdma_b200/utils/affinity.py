"""Pin the calling process to the CPUs (hence NUMA node) local to a GPU.

Pinned staging buffers are first-touched after this call, so H2D/D2H traffic stays on the
GPU's own PCIe root complex -- the reference's only performance advice is exactly this
topology rule ("GPU and HCA on the same PCIe root complex", README.md:71-72)."""
from __future__ import annotations

import os


def gpu_cpu_affinity(device: int):
    """Return the sorted list of CPU ids NVML reports as local to CUDA device ``device``."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device)
        bus = f"{getattr(p, 'pci_domain_id', 0):08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode() if isinstance(bus, str) else bus)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, v in enumerate(words) for b in range(64) if (v >> b) & 1]
        return [c for c in cpus if c < ncpu]
    except Exception:
        return []


def bind_to_gpu(device: int) -> list:
    cpus = gpu_cpu_affinity(device)
    if cpus:
        try:
            allowed = os.sched_getaffinity(0)
            want = set(cpus) & allowed
            if want:
                os.sched_setaffinity(0, want)
                return sorted(want)
        except OSError:
            pass
    return []
