"""Batch sharding of clips across the GPUs of one box.

Clips are independent in eval mode (BatchNorm uses running statistics), so the batch is split
into contiguous clip ranges, one process per GPU, with NO collective on the data path; the only
exchange is one all-gather of the (clips, 35) result rows -- 12 key probabilities, 12 tonic
logits, 11 genre logits -- and, optionally, one all-reduce of a few metric counters.  The
reference itself is single-GPU (train_model.py:86,116); this layer is new.

Works with the ``nccl`` backend on CUDA tensors (production) and with ``gloo`` on CPU tensors
(host-logic tests, tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

ROW = 35  # 12 key + 12 tonic + 11 genre


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) clip range of ``rank``; the first ``n_clips % world`` ranks hold one extra clip."""
    if world <= 0 or not 0 <= rank < world or n_clips < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(n_clips: int, world: int) -> List[int]:
    return [shard_range(n_clips, r, world)[1] - shard_range(n_clips, r, world)[0] for r in range(world)]


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the process group if world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host buffers:
    first-touch then places those buffers in the memory next to the GPU's PCIe root, so N ranks streaming audio to N GPUs do
    not all pull from one socket's memory.  Best effort: returns the node, or None when the topology is not exposed (single
    node, container without sysfs) -- nothing is changed then."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # best effort by contract: no CUDA, no sysfs, odd cpulist -> leave the affinity alone
        return None


def pack_rows(key: torch.Tensor, tonic: torch.Tensor, genre: Optional[torch.Tensor]) -> torch.Tensor:
    """(n, 35) fp32 rows [key | tonic | genre-or-zeros]."""
    n = key.shape[0]
    rows = torch.zeros((n, ROW), dtype=torch.float32, device=key.device)
    rows[:, :12] = key
    rows[:, 12:24] = tonic
    if genre is not None:
        rows[:, 24:] = genre
    return rows


def unpack_rows(rows: torch.Tensor, genre: bool):
    out = (rows[:, :12], rows[:, 12:24])
    return out + ((rows[:, 24:],) if genre else ())


def gather_rows(local_rows: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """All ranks receive the (n_clips, 35) table in clip order.  One collective (all_gather_into_tensor on
    equal-size padded shards: NCCL's fast path; ragged tails are trimmed afterwards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local_rows.shape[0] != n_clips:
            raise ValueError("single process must hold every clip")
        return local_rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = shard_counts(n_clips, world)
    if local_rows.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {counts[rank]}")
    width = local_rows.shape[1]
    cmax = max(counts)
    send = local_rows
    if counts[rank] != cmax:
        send = torch.zeros((cmax, width), dtype=local_rows.dtype, device=local_rows.device)
        send[: counts[rank]] = local_rows
    recv = torch.empty((world * cmax, width), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(c == cmax for c in counts):
        return recv
    return torch.cat([recv[r * cmax: r * cmax + counts[r]] for r in range(world)], dim=0)


class RowGather:
    """The all-gather of one step's (clips, 35) result rows, issued asynchronously: NCCL runs it on the process group's own
    stream, ordered after the kernels already queued on the current stream, so it overlaps the NEXT step's kernels; ``wait()``
    makes the current stream wait for it and returns the (n_clips, 35) table in clip order.  (gloo: same API on CPU tensors.)"""

    def __init__(self, local_rows: torch.Tensor, n_clips: int, group=None):
        self._n, self._group = n_clips, group
        self._work, self._recv, self._counts = None, None, None
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            if local_rows.shape[0] != n_clips:
                raise ValueError("single process must hold every clip")
            self._recv = local_rows
            return
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        counts = shard_counts(n_clips, world)
        if local_rows.shape[0] != counts[rank]:
            raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {counts[rank]}")
        cmax, width = max(counts), local_rows.shape[1]
        send = local_rows.contiguous()
        if counts[rank] != cmax:
            send = torch.zeros((cmax, width), dtype=local_rows.dtype, device=local_rows.device)
            send[: counts[rank]] = local_rows
        self._recv = torch.empty((world * cmax, width), dtype=local_rows.dtype, device=local_rows.device)
        self._counts = counts
        self._work = dist.all_gather_into_tensor(self._recv, send, group=group, async_op=True)

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._counts is None or all(c == max(self._counts) for c in self._counts):
            return self._recv
        cmax = max(self._counts)
        return torch.cat([self._recv[r * cmax: r * cmax + c] for r, c in enumerate(self._counts)], dim=0)


def reduce_counters(counters: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a small vector of metric counters over ranks (in place)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def allreduce_gradients(flat_grads: torch.Tensor, group=None) -> torch.Tensor:
    """Average the flat gradient buffer of training.TrainStep over the ranks, in place: ONE all-reduce of
    162,902 (167,031 with the genre head) fp32 values per step -- latency-bound on NVLink/NVSwitch, so a single
    bucket (no per-tensor launches).  BatchNorm statistics stay per rank (plain DDP semantics)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
        flat_grads.div_(dist.get_world_size(group))
    return flat_grads
