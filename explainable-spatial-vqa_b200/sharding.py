"""Multi-GPU plumbing for the executor path: one process per GPU, questions sharded by contiguous ranges,
weights replicated, NO per-step collective - one gather of the results at the end.

Works with any torch.distributed backend: `nccl` on the GPU box (NVLink 5 / NVSwitch), `gloo` in the CPU
test-suite.  Every rank must call the collectives in the same order.
"""
from __future__ import annotations

import datetime
import os
from typing import Sequence

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first `n_items % world_size` ranks get one extra item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_ranges(costs: Sequence[int], world_size: int) -> list[tuple[int, int]]:
    """Contiguous split of items with per-item cost (FA: program steps per question) so every rank gets about
    the same total cost.  Order is preserved so results concatenate back in input order."""
    total = float(sum(costs))
    out, lo, acc = [], 0, 0.0
    n = len(costs)
    for r in range(world_size):
        target = total * (r + 1) / world_size
        hi = lo
        while hi < n and (acc + costs[hi] <= target or hi == lo) and (n - hi) > (world_size - 1 - r):
            acc += costs[hi]
            hi += 1
        if r == world_size - 1:
            hi = n
        out.append((lo, hi))
        lo = hi
    return out


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialises torch.distributed from torchrun's environment; returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        # a rank that falls out of step must fail fast, not hold N GPUs for the default 10-minute watchdog
        timeout = datetime.timedelta(seconds=int(os.environ.get("B200VQA_DIST_TIMEOUT_S", "120")))
        dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=timeout, **kwargs)
    return rank, local, world


def _parse_cpulist(text: str) -> list[int]:
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pins this process (and therefore its first-touch pinned host allocations) to the CPUs of the NUMA node its GPU
    hangs off, so that every rank's host->device copies read node-local memory instead of crossing the socket link.
    Returns what was found / done; a box that reports no topology (numa_node -1, one node) is left alone."""
    info = {"numa_node": None, "bound": False}
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["nodes"] = len(nodes)
        if node < 0 or len(nodes) < 2:
            return info
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
            info["cpus"] = len(allowed)
    except Exception as e:  # pragma: no cover - topology files differ between hosts; never fatal
        info["error"] = repr(e)
    return info


def gather_varlen(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gathers row blocks of different lengths (counts[r] rows from rank r) into one tensor ordered by
    rank - the single collective of the executor path (answers, programs or step caches)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {counts[rank]}")
    width = max(counts)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded.contiguous(), group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


class JobGather:
    """Results of a whole job - `steps` batches of `rows` result rows per rank - collected with ONE collective at the
    end: every step copies its rows into a preallocated device buffer (no communication), `finish()` runs
    a single all_gather_into_tensor into a preallocated [world, steps, rows, width] buffer."""

    def __init__(self, steps: int, rows: int, width: int, dtype, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = torch.zeros(steps, rows, width, dtype=dtype, device=device)
        self.full = torch.zeros(self.world, steps, rows, width, dtype=dtype, device=device) if self.world > 1 else None
        self.steps = steps

    def put(self, step: int, rows: torch.Tensor) -> None:
        """Device-to-device copy on the CURRENT stream (the stream that produced `rows`)."""
        self.local[step % self.steps].copy_(rows, non_blocking=True)

    def finish(self) -> torch.Tensor:
        """[world, steps, rows, width] on every rank; call after the producing streams have been joined."""
        if self.world == 1:
            return self.local[None]
        # concatenation form (output = the ranks' blocks along dim 0): accepted by NCCL and gloo alike
        dist.all_gather_into_tensor(self.full.view(-1, *self.local.shape[1:]), self.local, group=self.group)
        return self.full


def max_over_ranks(value: float, device=None) -> float:
    """Max-reduction of a scalar timing across ranks (the bench's max-over-ranks rule)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
