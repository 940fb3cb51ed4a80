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
    # ONE collective into ONE preallocated tensor (rank-major), no per-rank list of receive buffers
    recv = torch.empty(w * send.numel(), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send.reshape(-1))
    recv = recv.reshape((w,) + tuple(send.shape))
    full = torch.view_as_complex(recv) if is_complex else recv  # (w, ..., max_w)
    if full.ndim == 2 and num_columns == w * max_w:  # 1-d observables, even split: already in column order
        return full.reshape(-1)
    if num_columns == w * max_w:
        return torch.movedim(full, 0, -2).reshape(full.shape[1:-1] + (num_columns,))
    return torch.cat([full[r][..., : hi - lo] for r, (lo, hi) in enumerate(widths)], dim=-1)


def solve_lmde_sharded(generator, t_span, y0: torch.Tensor, gather: bool = True, **kwargs):
    """solve_lmde on this rank's column block of y0 (n, B); optionally all-gather results.y."""
    from .solvers import solve_lmde

    B = y0.shape[-1]
    res = solve_lmde(generator, t_span, shard_columns(y0), **kwargs)
    if gather:
        res.y = all_gather_columns(res.y, B)
    return res


def shard_list(items: List, rank: Optional[int] = None, world_size: Optional[int] = None) -> List:
    """This rank's contiguous block of a list of simulations (same split as :func:`shard_bounds`)."""
    lo, hi = shard_bounds(len(items), rank, world_size)
    return list(items[lo:hi])


SWEEP_SUB_BLOCK = 8192  # simulations per Solver.solve call of a sharded sweep


def solver_solve_sharded(solver, t_span, y0, signals: List, measurement=None, gather: bool = True, **kwargs):
    """``Solver.solve`` for a LIST of simulations -- a parameter sweep (BASELINE.json configs[4]: 65 536 points over the
    8 GPUs of a box) -- with the list split into contiguous blocks over the ranks.

    Every rank runs its block through ``solver.solve`` (one sweep-mode launch per block when the simulations qualify,
    solvers/solver_classes.py:556-590 semantics otherwise); nothing is exchanged while stepping.  Afterwards ONE
    all-gather moves either the final states ``(n, N)`` or, when ``measurement`` (a
    :class:`~qiskit_dynamics_b200.measurement.FinalStateMeasurement`) is given, the memory-slot outcome probabilities
    ``(n_out, N)`` -- the "final observables" of the north star.

    ``signals`` is the list of per-simulation signal specifications; ``y0`` is one initial state shared by all
    simulations or a list with one per simulation.  Returns ``(local_results, gathered)``: this rank's list of
    ``OdeResult`` and the gathered table (``None`` when ``gather`` is false; this rank's block when not distributed).
    """
    nsim = len(signals)
    local_signals = shard_list(signals)
    if isinstance(y0, list):
        if len(y0) != nsim:
            raise ValueError(f"y0 lists {len(y0)} initial states for {nsim} simulations")
        local_y0 = shard_list(y0)
    else:
        local_y0 = y0
    if not local_signals:  # more ranks than simulations: this rank idles but still joins the collective
        local_results = []
    else:
        # sub-blocks: Solver.solve only ENQUEUES the device work of a block, so the host compiles the signal lists of the next
        # block while the GPU integrates the previous one (65 536 points: the host side is 3/4 of the wall time)
        n_local = len(local_signals)
        local_results, finals_parts = None, []
        for lo in range(0, n_local, SWEEP_SUB_BLOCK):
            hi = min(n_local, lo + SWEEP_SUB_BLOCK)
            part = solver.solve(t_span=t_span, y0=local_y0[lo:hi] if isinstance(local_y0, list) else local_y0,
                                signals=local_signals[lo:hi], **kwargs)
            if not isinstance(part, list):
                part = [part]
            finals_parts.append(getattr(part, "final_states", None))
            if local_results is None:
                local_results = part
            else:
                local_results.extend(part)
        if len(finals_parts) > 1 and hasattr(local_results, "final_states"):
            local_results.final_states = (torch.cat(finals_parts, dim=-1) if all(f is not None for f in finals_parts) else None)
    if not gather:
        return local_results, None
    if local_results:
        finals = getattr(local_results, "final_states", None)  # a batched sweep hands over its (n, n_local) final states
        if finals is None:
            finals = torch.stack([r.y[-1] for r in local_results], dim=-1)  # (n, n_local)
        t_final = float(local_results[0].t[-1])
        table = measurement.probabilities(t_final, finals) if measurement is not None else finals
    else:
        table = None
    rank, w = world()
    if w > 1:
        # ranks without simulations need the table's row count and dtype to take part in the gather
        rows = torch.tensor([0 if table is None else table.shape[0], 0 if table is None else int(table.is_complex())],
                            dtype=torch.int64, device=_collective_device(table))
        dist.all_reduce(rows, op=dist.ReduceOp.MAX)
        if table is None:
            table = torch.zeros((int(rows[0]), 0), dtype=torch.complex128 if int(rows[1]) else torch.float64,
                                device=rows.device)
    return local_results, (all_gather_columns(table, nsim) if table is not None else None)


def _collective_device(t: Optional[torch.Tensor]) -> torch.device:
    if t is not None:
        return t.device
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")
