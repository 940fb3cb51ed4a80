"""Multi-GPU: the batch axis shards embarrassingly.

One process per GPU (torchrun), operators/frames/signal tables replicated, state columns split
into contiguous blocks, no communication while stepping, and ONE collective at the end: an
all-gather of the final states (or of per-column observables) over NCCL / NVLink.  With the
``gloo`` backend the same code runs on CPU tensors, which is how the sharding logic is tested
without GPUs.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(num_columns: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of columns owned by `rank`; earlier ranks take the remainder."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(num_columns, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_columns(y: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    """This rank's block of the trailing (column) axis, contiguous."""
    lo, hi = shard_bounds(y.shape[-1], rank, world_size)
    return y[..., lo:hi].contiguous()


def all_gather_columns(local: torch.Tensor, num_columns: int) -> torch.Tensor:
    """Gather per-rank column blocks (possibly ragged) back into the full trailing axis.
    Complex tensors travel as their real view (NCCL has no complex dtype)."""
    rank, w = world()
    if w == 1:
        return local
    widths = [shard_bounds(num_columns, r, w) for r in range(w)]
    max_w = max(hi - lo for lo, hi in widths)
    is_complex = local.is_complex()
    buf = local
    if buf.shape[-1] < max_w:  # pad ragged shards so that every rank sends the same size
        pad = torch.zeros(buf.shape[:-1] + (max_w - buf.shape[-1],), dtype=buf.dtype, device=buf.device)
        buf = torch.cat([buf, pad], dim=-1)
    send = torch.view_as_real(buf.contiguous()) if is_complex else buf.contiguous()
    recv = [torch.empty_like(send) for _ in range(w)]
    dist.all_gather(recv, send)
    parts = []
    for r, (lo, hi) in enumerate(widths):
        part = torch.view_as_complex(recv[r]) if is_complex else recv[r]
        parts.append(part[..., : hi - lo])
    return torch.cat(parts, dim=-1)


def solve_lmde_sharded(generator, t_span, y0: torch.Tensor, gather: bool = True, **kwargs):
    """solve_lmde on this rank's column block of y0 (n, B); optionally all-gather results.y."""
    from .solvers import solve_lmde

    B = y0.shape[-1]
    res = solve_lmde(generator, t_span, shard_columns(y0), **kwargs)
    if gather:
        res.y = all_gather_columns(res.y, B)
    return res
