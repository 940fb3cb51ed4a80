"""Array plumbing: everything numeric is a torch.complex128 (or float64) tensor on one device.

The reference's arraylias multi-backend dispatch (arraylias/alias.py) is removed: there is one
array type and one compute path.  `default_device()` is cuda:LOCAL_RANK when a GPU is visible;
without one, tensors can still be *constructed* (host logic, argument checks) but every
evaluation raises, because the CUDA C-ABI is the only implementation.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

CDTYPE = torch.complex128
RDTYPE = torch.float64

_device_override: Optional[torch.device] = None


def set_default_device(device) -> None:
    global _device_override
    _device_override = None if device is None else torch.device(device)


def default_device() -> torch.device:
    if _device_override is not None:
        return _device_override
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1))
    return torch.device("cpu")


_pending_copies = []


def wait_pending_copies() -> None:
    """Block until every asynchronous pinned-host -> device copy issued by :func:`asarray` has completed.  The
    solve entry points call this before returning, so a caller may overwrite its pinned input buffer as soon as
    the solve call is back (the copy is long finished by then; the wait costs nothing)."""
    while _pending_copies:
        _pending_copies.pop().synchronize()


_pinned_staging = {}
_staging_event = {}


def stage_to_device(values: np.ndarray, device) -> torch.Tensor:
    """Small float64 host array -> device through a cached pinned staging buffer, without blocking the host
    (``torch.from_numpy(x).to(device)`` from pageable memory synchronises the stream).  The buffer is reused by the
    next call, which is safe because the solve entry points wait for the copy event before they return."""
    device = torch.device(device)
    values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
    if device.type != "cuda":
        return torch.from_numpy(values.copy()).to(device)
    prev = _staging_event.get(device)  # the staging buffer may still be the source of the previous grid's copy
    if prev is not None:
        prev.synchronize()
    buf = _pinned_staging.get(device)
    if buf is None or buf.numel() < values.size:
        buf = torch.empty(max(values.size, 4096), dtype=torch.float64).pin_memory()
        _pinned_staging[device] = buf
    buf[: values.size].copy_(torch.from_numpy(values))
    out = buf[: values.size].to(device, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    _staging_event[device] = ev
    _pending_copies.append(ev)
    return out


def asarray(x, device=None) -> Optional[torch.Tensor]:
    """Anything array-like (numpy, nested lists, objects with __array__/.data, tensors) ->
    contiguous complex128 tensor on the compute device.  None passes through."""
    if x is None:
        return None
    device = default_device() if device is None else torch.device(device)
    if isinstance(x, torch.Tensor):
        # resolve lazy conj/neg bits: raw device pointers cross the C-ABI.  A pinned host tensor is copied
        # asynchronously on the current stream, so that the host goes on to build the step grid and the signal
        # table while the state batch is in flight (every consumer is enqueued on the same stream).
        nb = x.device.type == "cpu" and device.type == "cuda" and x.is_pinned()
        out = x.to(device=device, dtype=CDTYPE, non_blocking=nb).resolve_conj().resolve_neg().contiguous()
        if nb:  # the caller's buffer must not be reused before the copy has landed: see wait_pending_copies
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            _pending_copies.append(ev)
        return out
    if isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], torch.Tensor):
        return torch.stack([asarray(e, device) for e in x]).contiguous()
    arr = np.asarray(x, dtype=complex) if not (isinstance(x, (list, tuple)) and len(x) and hasattr(x[0], "data") and not isinstance(x[0], np.ndarray)) \
        else np.asarray([np.asarray(getattr(e, "data", e)) for e in x], dtype=complex)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device=device, dtype=CDTYPE)


def asreal(x, device=None) -> torch.Tensor:
    device = default_device() if device is None else torch.device(device)
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=RDTYPE).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device)


def to_numpy(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)
