"""Host-side placement: keep a rank's threads and its page-locked buffers on the NUMA node next to its GPU.

The host pipeline is PCIe-bound; with 8 ranks on a two-socket box every rank's D2H stream
should land in the memory of the socket its GPU hangs off, otherwise half of the traffic crosses the inter-socket
link.  Linux allocates pages on the node of the CPU that touches them first, so pinning the process to the GPU's
CPU set BEFORE allocating the pinned buffers is enough.  Best effort: without NVML (or on a single-node box)
this is a no-op.
"""
import os


def gpu_cpu_set(device=0):
    """CPUs NVML reports as local to `device` (empty set when unknown)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        n_cpu = os.cpu_count() or 1
        words = (n_cpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        return {c for c in cpus if c < n_cpu}
    except Exception:
        return set()


def bind_host_to_gpu(device=0):
    """Restrict this process to the CPUs local to `device`; returns the CPU set in effect (possibly unchanged)."""
    cpus = gpu_cpu_set(device)
    try:
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
        return os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return set()
