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


class PeerGather:
    """All-gather of equal-size shards by peer-to-peer copies over NVLink on the COPY ENGINES: the destination buffers
    are symmetric memory (torch.distributed._symmetric_memory: cuMem allocations exported to every rank of the node),
    every rank pushes its shard into all of them with plain device-to-device copies (measured 715 GB/s per push on
    B200 / NV18), and a 4-byte all-reduce behind the pushes is the completion barrier.  An NCCL all-gather of the same
    bytes runs as a kernel on a dozen SMs for its whole duration, and this path's compute kernels are persistent with
    one CTA per SM: whenever the gather of step i overlaps the kernels of step i + 1, those kernels wait for the SMs
    NCCL holds.  Copy-engine pushes take no SM at all.  (Legacy cudaIpc mappings — torch.multiprocessing's
    reduce_tensor — were measured first: 25 GB/s per push across processes in this environment, i.e. staged through
    the host; not used.)

    `slots` destination buffers alternate between calls (the gather of one step overlaps the next step's compute).
    `gather(shard, slot, stream, consumed=None)`: enqueue on `stream`; returns this rank's (world * shard_numel,) buffer
    of that slot, valid once `stream` has run past the call.  `consumed`: event recorded after this rank finished
    READING that slot the last time round — peers overwrite it two calls later, and the barrier of the call in between
    is what holds them back until every rank has passed its `consumed` event."""

    def __init__(self, shard_numel: int, dtype, device, slots: int = 2, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group, self.n, self.slots = group, int(shard_numel), slots
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        name = (group or dist.group.WORLD).group_name
        self.bufs, self.peer, err = [], [[] for _ in range(self.world)], None
        try:
            for _ in range(slots):
                buf = symm_mem.empty(self.world * self.n, dtype=dtype, device=self.device)
                hdl = symm_mem.rendezvous(buf, name)
                self.bufs.append(buf)
                for r in range(self.world):
                    self.peer[r].append(buf if r == self.rank else hdl.get_buffer(r, (self.world * self.n,), dtype))
        except Exception as e:             # no symmetric memory here: fail on EVERY rank, not on one
            err = e
        self._token = torch.zeros(1, device=self.device)
        ok = torch.tensor([0 if err else 1], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if not ok.item():
            raise RuntimeError(f"PeerGather: symmetric memory is not available on every rank ({err})")

    def gather(self, shard: torch.Tensor, slot: int, stream=None, consumed=None) -> torch.Tensor:
        if shard.numel() != self.n or shard.dtype != self.bufs[0].dtype or not shard.is_contiguous():
            raise RuntimeError("shard must be contiguous with the size and dtype given at construction")
        stream = stream or torch.cuda.current_stream(self.device)
        flat = shard.reshape(-1)
        with torch.cuda.stream(stream):
            if consumed is not None:
                stream.wait_event(consumed)
            for k in range(self.world):
                r = (self.rank + k) % self.world                      # spread the targets: no two ranks start on the same peer
                self.peer[r][slot][self.rank * self.n:(self.rank + 1) * self.n].copy_(flat, non_blocking=True)
            dist.all_reduce(self._token, group=self.group)             # nobody gets past it before everybody's pushes are done
        return self.bufs[slot]


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

