"""Multi-GPU sharding of the zoom-FFT PSD path (one process per GPU).

Frames are independent -- the LO phase, the filter state and the Welch mean
all restart per chunk (pypanadapter_spectrum.py:2092, 2098, 2111) -- and so
are receiver channels, so the path shards with NO data-path collective: rank
r processes its contiguous block of frames (or its channels) on its own GPU.
The only exchange is the gather of finished rows (W float32 each) to rank 0,
over NCCL/NVLink on GPUs (gloo in the CPU tests).  EMA is order dependent:
with ``ema_alpha`` set, every rank averages over its own block of frames
(= its own receiver stream); shard by channel, not by time, if one stream's
EMA has to span all frames.
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``n`` units for ``rank``; blocks differ by
    at most one unit and concatenate to range(n) in rank order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_rows(local_rows, total_units: int, *, dst: int = 0, group=None):
    """Gather every rank's rows to ``dst`` in rank order.

    ``local_rows``: torch tensor (n_local, W) float32 on the device the
    process group's backend works with (CUDA for nccl, CPU for gloo).
    Returns (total_units, W) on ``dst``, None elsewhere.  Ragged blocks are
    padded to the largest block for the collective and trimmed afterwards.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = shard_counts(total_units, world)
    if local_rows.shape[0] != counts[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local_rows.shape[0], counts[rank]))
    width = local_rows.shape[1]
    most = max(counts)
    send = local_rows
    if counts[rank] != most:
        send = torch.zeros((most, width), dtype=local_rows.dtype, device=local_rows.device)
        send[:counts[rank]] = local_rows
    send = send.contiguous()
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:counts[r]] for r in range(world)], dim=0)


def process_frames_sharded(engine, frames: np.ndarray, *, dst: int = 0, group=None, device=None):
    """All ranks hold (or can index) the same ``frames`` array; each processes
    its block through ``engine`` (already configured) and rank ``dst`` gets
    all rows in frame order -- bit-identical to a single-GPU run without EMA."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = len(frames)
    lo, hi = shard_range(n, rank, world)
    if hi > lo:
        rows = engine.process(frames[lo:hi])
    else:
        rows = np.empty((0, engine.row_width), dtype=np.float32)
    t = torch.from_numpy(rows)
    if device is not None:
        t = t.to(device)
    out = gather_rows(t, n, dst=dst, group=group)
    return None if out is None else out.cpu().numpy()


def channels_for_rank(nchannels: int, rank: int, world: int) -> range:
    """Receiver channels (distinct zoom centres over one stream, BASELINE
    configs[3]) owned by ``rank``."""
    lo, hi = shard_range(nchannels, rank, world)
    return range(lo, hi)


# ---------------------------------------------------------------------------
# host placement: one process per GPU, each next to its GPU
# ---------------------------------------------------------------------------
def parse_cpulist(text: str) -> set[int]:
    """Linux cpulist syntax ("0-31,64-95") -> set of CPU numbers."""
    cpus: set[int] = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(pci_bus_id: str, sysfs: str = "/sys") -> set[int]:
    """CPUs of the NUMA node the GPU at ``pci_bus_id`` ("0000:1b:00.0", any case,
    8-digit nvml domains accepted) hangs off, from sysfs; empty if unknown."""
    bus = pci_bus_id.strip().lower()
    dom, _, rest = bus.partition(":")
    if len(dom) > 4:
        bus = dom[-4:] + ":" + rest
    try:
        with open(os.path.join(sysfs, "bus", "pci", "devices", bus, "local_cpulist")) as f:
            return parse_cpulist(f.read())
    except (OSError, ValueError):
        return set()


def bind_host_to_gpu(pci_bus_id: str, sysfs: str = "/sys") -> dict:
    """Restrict this process to the CPUs local to its GPU, so that the pinned
    staging memory it allocates afterwards (first touch) and the threads that
    feed the copy engine sit on the GPU's own NUMA node: with 8 ranks each
    streaming ~53 GB/s of samples over PCIe, buffers that land on the other
    socket cross the inter-socket link twice.  Never fails: returns
    ``{"bound": False, "why": ...}`` when sysfs has no answer or the CPUs are
    outside this process's cpuset."""
    try:
        allowed = os.sched_getaffinity(0)
    except (AttributeError, OSError) as exc:
        return {"bound": False, "why": "no sched_getaffinity: %s" % exc}
    local = gpu_local_cpus(pci_bus_id, sysfs)
    if not local:
        return {"bound": False, "why": "no local_cpulist for %s" % pci_bus_id}
    want = local & allowed
    if not want:
        return {"bound": False, "why": "local CPUs of %s are outside the cpuset" % pci_bus_id}
    if want == allowed:
        return {"bound": False, "why": "single NUMA domain", "cpus": len(allowed)}
    try:
        os.sched_setaffinity(0, want)
    except OSError as exc:
        return {"bound": False, "why": "sched_setaffinity: %s" % exc}
    return {"bound": True, "cpus": len(want), "of": len(allowed)}
