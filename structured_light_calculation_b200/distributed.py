"""Multi-GPU plumbing: one process per GPU, frame sets sharded, no data-path collective.

Every frame set (stack) of a DynaFrame sequence is independent in the north-star
definition, so the only cross-rank traffic is the barrier and the max-over-ranks
of the measured time (torch.distributed, NCCL on GPUs / gloo on CPU in tests).
Calibration is replicated per rank at context creation.
"""
from __future__ import annotations

import os


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of frame-set indices for `rank`; sizes differ by <= 1."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request n={n_items} rank={rank} world={world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment, (0, 0, 1) if absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str):
    import torch.distributed as dist
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        kw = {}
        if backend == "nccl":
            import torch
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_local_cpus(pci_domain: int, pci_bus: int, pci_device: int) -> set[int]:
    """CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist); empty if unknown."""
    path = f"/sys/bus/pci/devices/{pci_domain:04x}:{pci_bus:02x}:{pci_device:02x}.0/local_cpulist"
    try:
        with open(path) as f:
            return _parse_cpulist(f.read())
    except OSError:
        return set()


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process to the cores next to its GPU, so that the pinned staging buffers
    it allocates afterwards are first-touched on that NUMA node and the H2D/D2H copies do
    not cross the socket interconnect (SURVEY hard part 8).  Returns what was done."""
    import torch
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(local_rank)
        cpus = gpu_local_cpus(props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        info.update(pci=f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0",
                    local_cpus=len(cpus), allowed=len(allowed))
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info.update(bound=True, cpus=len(target))
    except Exception as e:  # best effort: never fail a run because of placement
        info["error"] = repr(e)
    return info
