"""Image-sharded data-parallel decode: rank r of N decodes images
{i : i mod N == r}; weights are replicated; there is NO collective on the data path.  The only
communication is one gather of a few scalars per rank after the decode loop (NCCL on GPUs, gloo
in the CPU tests).  One process per GPU (torchrun)."""
import os
import time
from typing import Callable, Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_images: int, rank: int, world: int) -> List[int]:
    """Images of rank `rank`: i = rank, rank + world, ...  Every image belongs to exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_images, world))


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def decode_sharded(decode_fn: Callable[[int], torch.Tensor], n_images: int, rank: int, world: int,
                   on_result: Callable[[int, torch.Tensor], None] = None):
    """Run decode_fn(i) for this rank's images.  Returns (n_done, seconds)."""
    t0 = time.perf_counter()
    n = 0
    for i in shard_indices(n_images, rank, world):
        out = decode_fn(i)
        if on_result is not None:
            on_result(i, out)
        n += 1
    if torch.cuda.is_available() and torch.cuda.is_initialized():
        torch.cuda.synchronize()
    return n, time.perf_counter() - t0


def gather_metrics(values: Sequence[float], device="cpu"):
    """All ranks contribute a small vector of scalars; every rank gets the [world, len] table back
    (one all_gather, outside any timed region)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None, :].cpu()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu()


def aggregate_throughput(table: torch.Tensor) -> float:
    """table rows = (n_images, seconds): whole-job images/s = total images / slowest rank's time."""
    return float(table[:, 0].sum() / table[:, 1].max())
