"""Clip sharding across the GPUs of one box (one process per GPU, torch.distributed).

With the plain generator (Full_model/Models.py) every clip is independent in eval mode (BatchNorm uses
running statistics; InstanceNorm, SE, LayerNorm and attention are per clip), so the forward needs no
collective and 1-GPU and N-GPU poses are bit-identical.  The Models_memory generator is the exception by
construction of the reference: its temporal memory sums over the clips of one call
(Full_model/Models_memory.py:287-288), so there a shard behaves like one nn.DataParallel replica of the
reference — clips are coupled within a rank's shard, not across ranks — and results depend on the
partition exactly as they do in the reference.  Communication is
exactly the two steps BASELINE.json's north star names: the final pose gather and the FGD
sufficient-statistics all-reduce (fgd.all_reduce_stats).  This replaces the reference's
single-process nn.DataParallel scatter/replicate/gather per call
(test_emotion_gesture_diversity_iterative.py:137-138).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_clips: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by <= 1)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(n_clips, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_poses(local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """Gather per-rank pose slices (b_r, F, P) into the global (n_clips, F, P) on every rank.
    Shards may be ragged (n_clips not divisible by the world size): slices are padded to the
    largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_clips
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_clips, r, world) for r in range(world)]
    assert local.shape[0] == sizes[rank][1] - sizes[rank][0], "local shard has the wrong size"
    biggest = max(hi - lo for lo, hi in sizes)
    send = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        send = torch.cat([local, pad], dim=0)
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    parts = [out[r * biggest: r * biggest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def bind_host_to_gpu_node(device_index: int):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers allocated
    afterwards (first-touch placement) sit next to that GPU's PCIe root: with one process per GPU streaming ~600 MB per
    step each, cross-socket copies otherwise share the inter-socket link.  Best effort: returns the node number, or
    None when the topology cannot be read (containers without /sys PCI entries, single-node hosts) — never raises."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev_id = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None

