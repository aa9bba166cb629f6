"""Host staging memory placed next to the GPU that will read / write it.

`CLIP.predict_host` moves ~0.8 GB per 50k x 1000 step over PCIe.  With one process per GPU on a two-socket host, pinned
buffers that land on the other socket send every byte across the inter-socket link as well; at 8 ranks that link, not
PCIe, bounds the end-to-end rate.  Page-locked memory is physically allocated by the thread that requests it, so it is
enough to run the allocating thread on the CPUs NVML reports as local to the device while the buffer is created.
Everything here is best effort: without NVML (or on a single-node host) the helpers degrade to plain pinned allocation.
"""
import contextlib
import os
from typing import Optional, Set

import torch


def gpu_local_cpus(device) -> Optional[Set[int]]:
    """CPUs of the NUMA node the GPU hangs off (NVML ideal CPU affinity), restricted to this process' allowed set."""
    try:
        import pynvml

        dev = torch.device(device)
        props = torch.cuda.get_device_properties(dev)
        pynvml.nvmlInit()
        bus_id = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(os.cpu_count() or 1, 1) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:  # no NVML / unsupported platform: leave placement to the OS
        return None


@contextlib.contextmanager
def numa_local(device):
    """Run the calling thread on the GPU-local CPUs for the duration of the block (memory it allocates is node local)."""
    cpus = gpu_local_cpus(device) if torch.device(device).type == "cuda" else None
    if not cpus or not hasattr(os, "sched_setaffinity"):
        yield False
        return
    before = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
        yield True
    finally:
        os.sched_setaffinity(0, before)


def pinned_empty(shape, dtype=torch.float32, device="cuda") -> torch.Tensor:
    """Page-locked host tensor allocated on the NUMA node of `device`."""
    with numa_local(device):
        return torch.empty(tuple(shape), dtype=dtype, pin_memory=True)


def pin(tensor: torch.Tensor, device="cuda") -> torch.Tensor:
    """Page-locked, GPU-local copy of a host tensor."""
    out = pinned_empty(tensor.shape, tensor.dtype, device)
    out.copy_(tensor)
    return out
