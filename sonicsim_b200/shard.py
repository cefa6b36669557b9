"""Multi-GPU sharding of the render path.

Unit of work = one (utterance, source) pair: independent dry signal, RIR set, trajectory and
output (SonicSet.py:77-94), so units shard across ranks with NO data-path collective.  The only
collective is one all-gather of a 3-float counter struct per rank at the end (NCCL on GPUs, gloo
in the CPU tests).  One process per GPU.
"""
import os
import typing as T

import numpy as np


def unit_cost(N: int, P: int, C: int, L: int, moving: bool) -> float:
    """Relative cost of a unit: inverse transforms executed (blocks x channels, ~x1.3 for the
    blocks that straddle a waypoint) plus the spectra of its RIRs and dry windows."""
    nb = (N + 4095) // 4096
    K = (L + 4095) // 4096
    if moving:
        return nb * C * (1.0 + min(1.0, (P - 1) / max(nb, 1))) * K + 0.5 * (P * C * K + nb)
    return nb * ((C + 1) // 2) * K + 0.5 * (C * K + nb)


def shard_units(costs: T.Sequence[float], world_size: int, rank: int) -> T.List[int]:
    """Sort by cost (moving >> static), deal round-robin in a serpentine so every rank gets the same
    count (+-1) and nearly the same total cost.  Deterministic; every rank computes the same plan."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    mine = []
    for pos, i in enumerate(order):
        rnd, k = divmod(pos, world_size)
        owner = k if rnd % 2 == 0 else world_size - 1 - k
        if owner == rank:
            mine.append(i)
    return sorted(mine)


def dist_env() -> T.Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def gather_counters(audio_seconds: float, elapsed_seconds: float, alg_bytes: float, device=None) -> np.ndarray:
    """All-gather (audio_seconds, elapsed_seconds, alg_bytes) from every rank -> (world, 3) array.
    This is the path's single collective."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([audio_seconds, elapsed_seconds, alg_bytes], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.cpu().numpy()[None, :]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def aggregate_throughput(counters: np.ndarray) -> T.Tuple[float, float]:
    """Whole-job audio-seconds per second = total audio / max elapsed over ranks; and total bytes/s."""
    total_audio = float(counters[:, 0].sum())
    t_max = float(counters[:, 1].max())
    return total_audio / t_max, float(counters[:, 2].sum()) / t_max


def bind_to_gpu_numa(cuda_index: int) -> T.Optional[int]:
    """Pin this process (and therefore its first-touch pinned host buffers) to the CPUs of the NUMA node
    the GPU hangs off, so that the H2D / D2H copies of the host path do not cross sockets.  Best effort:
    returns the node id or None when the topology cannot be read."""
    try:
        import torch
        props = torch.cuda.get_device_properties(cuda_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:          # noqa: BLE001
        return None
