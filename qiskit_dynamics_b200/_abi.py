"""ctypes binding of libqdb.so (include/qdb.h).  PyTorch is used only as the owner of device
memory and streams: every call passes raw device pointers + sizes, no torch types cross the ABI.

There is NO fallback: if the shared library is missing, or a tensor is not a CUDA complex128 /
float64 tensor, the call raises.
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libqdb.so")

LAYOUT_ROWMAJOR = 0
LAYOUT_PACKED = 1
LAYOUT_PACKED3M = 2
WS_RHS, WS_RK4, WS_EXPM, WS_MAGNUS, WS_PROP = 0, 1, 2, 3, 4


class QdbError(RuntimeError):
    """A libqdb call failed (message from qdb_last_error_string)."""


class c128(ctypes.Structure):
    _fields_ = [("re", ctypes.c_double), ("im", ctypes.c_double)]


_vp, _i, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t

# name -> (restype, argtypes); the single source of truth checked against include/qdb.h in tests
SIGNATURES = {
    "qdb_last_error_string": (ctypes.c_char_p, []),
    "qdb_version": (_i, []),
    "qdb_npad": (_i, [_i]),
    "qdb_packed_elems": (_sz, [_i]),
    "qdb_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "qdb_pack_operators": (_i, [_i, _i, _vp, _vp, _vp]),
    "qdb_generator_c128": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _d, _vp, _vp]),
    "qdb_frame_apply_c128": (_i, [_i, _i, _vp, _d, _i, _vp, _vp, _i, _vp]),
    "qdb_zgemm_c128": (_i, [_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, c128, c128, _vp, _vp, _vp, _vp]),
    "qdb_rhs_c128": (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _d, _vp, _vp, _i, _vp, _sz, _vp]),
    "qdb_rk4_steps_c128": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _d, _vp, _i, _vp, _sz, _vp]),
    "qdb_expm_steps_c128": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _i, _vp, _sz, _vp]),
    "qdb_rk4_table_steps_c128": (_i, [_i, _i, _i, _vp, _i, _d, _vp, _i, _vp]),
    "qdb_rk4_ozaki_workspace_bytes": (_sz, [_i]),
    "qdb_rk4_int8_preferred": (_i, [_i, _i]),
    "qdb_rk4_ozaki_slice_c128": (_i, [_i, _i, _vp, _i, _vp, _sz, _vp]),
    "qdb_rk4_ozaki_steps_c128": (_i, [_i, _i, _i, _vp, _d, _vp, _i, _vp, _sz, _vp]),
    "qdb_rk4_table_layout": (_i, [_i, _i]),
    "qdb_table_entry_bytes": (_sz, [_i, _i]),
    "qdb_rk4_tiling": (_i, [_i, _i, _i, _vp]),
    "qdb_signal_table_f64": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, ctypes.c_longlong, _vp, _vp, _d,
                                  _vp, _vp]),
    "qdb_outcome_probabilities_f64": (_i, [_i, _i, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "qdb_dmma_probe": (_i, [_vp, _i, _vp, _vp]),
    "qdb_expm_c128": (_i, [_i, _vp, _i, _vp, _vp, _sz, _vp]),
    "qdb_magnus_steps_c128": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _i, _vp, _sz, _vp]),
    "qdb_magnus_terms_c128": (_i, [_i, _i, _vp, _d, _d, _vp, _vp, _sz, _vp]),
    "qdb_step_propagators_c128": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp, _sz, _vp]),
    "qdb_lindblad_supported": (_i, [_i]),
    "qdb_lindblad_rhs_c128": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp, _vp]),
    "qdb_lindblad_rk4_steps_c128": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp]),
    "qdb_launch_count": (ctypes.c_ulonglong, []),
}

_lib = None


def lib():
    """Load libqdb.so (once).  Raises QdbError when it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QdbError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  qiskit_dynamics_b200 has no CPU fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().qdb_last_error_string().decode()
        raise QdbError(f"{what} failed (rc={rc}): {msg}")


def _ptr(t: Optional[torch.Tensor], dtype, name: str):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise QdbError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback), got {type(t).__name__}"
                       + (f" on {t.device}" if isinstance(t, torch.Tensor) else ""))
    if t.dtype != dtype:
        raise QdbError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise QdbError(f"{name}: tensor must be contiguous")
    if t.is_conj() or t.is_neg():
        raise QdbError(f"{name}: tensor carries a lazy conj/neg bit; call .resolve_conj()/.resolve_neg() first")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


C = torch.complex128
F = torch.float64


def npad(n: int) -> int:
    return (n + 7) & ~7


def packed_elems(n: int) -> int:
    """Elements of one operator in QDB_LAYOUT_PACKED (rows padded to 8, columns to 16)."""
    return ((n + 7) & ~7) * ((n + 15) & ~15)


def workspace_bytes(kind: int, n: int, K: int, B: int, S: int = 1) -> int:
    return int(lib().qdb_workspace_bytes(kind, n, K, B, S))


def launch_count() -> int:
    return int(lib().qdb_launch_count())


def rk4_tiling(n: int, B: int, sweep_K: int = 0) -> dict:
    """Tiling the on-chip RK4 kernel picks for (n, B) -- diagnostic (qdb_rk4_tiling)."""
    out = (ctypes.c_int * 9)()
    _check(lib().qdb_rk4_tiling(n, B, sweep_K, ctypes.cast(out, ctypes.c_void_p)), "qdb_rk4_tiling")
    keys = ("warps_rows", "warps_cols", "row_tiles_per_warp", "col_tiles_per_warp", "split", "ctas", "threads", "smem_bytes",
            "m3")
    return dict(zip(keys, (int(v) for v in out)))


def pack_operators(ops: torch.Tensor) -> torch.Tensor:
    """(count, n, n) row-major -> (count, npad*npad) DMMA-fragment order."""
    count, n, _ = ops.shape
    out = torch.empty((count, packed_elems(n)), dtype=C, device=ops.device)
    _check(lib().qdb_pack_operators(n, count, _ptr(ops, C, "ops"), _ptr(out, C, "out"), _stream()), "qdb_pack_operators")
    return out


def generator(n, ops, stat, coeff, mu, times, scale=1.0, layout=LAYOUT_ROWMAJOR, out=None):
    """out[t] = scale * (stat + sum_j coeff[t,j] ops[j]) .* outer(conj p(t), p(t)); coeff (T,K)."""
    K = 0 if ops is None else ops.shape[0]
    if coeff is not None:
        T = coeff.shape[0]
    elif times is not None:
        T = times.shape[0]
    else:
        T = 1
    cplx = coeff is not None and coeff.dtype == C
    elems = {LAYOUT_ROWMAJOR: n * n, LAYOUT_PACKED: packed_elems(n), LAYOUT_PACKED3M: packed_elems(n) * 3 // 2}[layout]
    dev = (ops if ops is not None else stat).device
    if out is None:
        out = torch.empty((T, elems), dtype=C, device=dev)
    _check(lib().qdb_generator_c128(n, K, T, layout, _ptr(ops, C, "ops"), _ptr(stat, C, "stat"),
                                    _ptr(coeff, C if cplx else F, "coeff"), int(cplx), _ptr(mu, F, "mu"),
                                    _ptr(times, F, "times"), float(scale), _ptr(out, C, "out"), _stream()),
           "qdb_generator_c128")
    return out


def frame_apply(mu, t, y, conj_phase: bool, out=None):
    """Rows of y (n, B) times exp(-i mu t) (or its conjugate)."""
    n, B = y.shape
    if out is None:
        out = torch.empty_like(y)
    _check(lib().qdb_frame_apply_c128(n, B, _ptr(mu, F, "mu"), float(t), int(conj_phase), _ptr(y, C, "y_in"),
                                      _ptr(out, C, "y_out"), B, _stream()), "qdb_frame_apply_c128")
    return out


def zgemm(A, Bm, out=None, alpha=1.0 + 0j, beta=0.0 + 0j, colscale=None, pre=None, post=None):
    M, Kd = A.shape
    Kd2, N = Bm.shape
    if Kd != Kd2:
        raise QdbError(f"zgemm: inner dimensions differ ({Kd} vs {Kd2})")
    if out is None:
        out = torch.empty((M, N), dtype=C, device=A.device)
    a, b = complex(alpha), complex(beta)
    _check(lib().qdb_zgemm_c128(M, N, Kd, _ptr(A, C, "A"), Kd, _ptr(Bm, C, "B"), N, _ptr(out, C, "C"), N,
                                c128(a.real, a.imag), c128(b.real, b.imag), _ptr(colscale, F, "colscale"),
                                _ptr(pre, C, "pre"), _ptr(post, C, "post"), _stream()), "qdb_zgemm_c128")
    return out


_rhs_ws_cache = {}


def rhs(n, ops, stat, coeff, mu, t, y, per_col=False, out=None, workspace=None):
    """One fused RHS evaluation on y (n, B)."""
    K = 0 if ops is None else ops.shape[0]
    B = y.shape[1]
    if out is None:
        out = torch.empty_like(y)
    if workspace is None:  # small (n^2 + 2n complex): keep one per (device, n, stream) instead of allocating per call
        key = (y.device, n, torch.cuda.current_stream(y.device).cuda_stream)  # per stream: concurrent streams must not share G(t)
        workspace = _rhs_ws_cache.get(key)
        if workspace is None:
            workspace = torch.empty(workspace_bytes(WS_RHS, n, K, B), dtype=torch.uint8, device=y.device)
            _rhs_ws_cache[key] = workspace
    need = workspace.numel() if workspace.numel() >= (n * n + 2 * n + 64) * 16 else workspace_bytes(WS_RHS, n, K, B)
    if workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=y.device)
    ldc = coeff.shape[-1] if (per_col and coeff is not None) else 0
    _check(lib().qdb_rhs_c128(n, K, B, _ptr(ops, C, "ops"), _ptr(stat, C, "stat"), _ptr(coeff, F, "coeff"),
                              int(per_col), ldc, _ptr(mu, F, "mu"), float(t), _ptr(y, C, "y_in"), _ptr(out, C, "y_out"),
                              B, ctypes.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()), "qdb_rhs_c128")
    return out


def rk4_steps(n, ops_rm, stat_rm, ops_packed, stat_packed, coeff, mu, times_host: np.ndarray, h, y, S,
              per_col=False, workspace=None, max_ws_bytes=1 << 30):
    """S fused RK4 steps in place on y (n, B).  times_host: float64 numpy array [2S+1]."""
    K = 0 if (ops_rm is None and ops_packed is None) else (ops_rm if ops_rm is not None else ops_packed).shape[0]
    B = y.shape[1]
    times_host = np.ascontiguousarray(times_host, dtype=np.float64)
    if times_host.shape[0] != 2 * S + 1:
        raise QdbError(f"rk4_steps: need {2 * S + 1} stage times, got {times_host.shape[0]}")
    if workspace is None:
        need = min(workspace_bytes(WS_RK4, n, K, B, S), max(max_ws_bytes, workspace_bytes(WS_RK4, n, K, B, 1)))
        workspace = torch.empty(need, dtype=torch.uint8, device=y.device)
    ldc = coeff.shape[-1] if (per_col and coeff is not None) else 0
    _check(lib().qdb_rk4_steps_c128(n, K, B, S, _ptr(ops_rm, C, "ops_rm"), _ptr(stat_rm, C, "stat_rm"),
                                    _ptr(ops_packed, C, "ops_packed"), _ptr(stat_packed, C, "stat_packed"),
                                    _ptr(coeff, F, "coeff"), int(per_col), ldc, _ptr(mu, F, "mu"),
                                    times_host.ctypes.data_as(ctypes.c_void_p), float(h), _ptr(y, C, "y"), B,
                                    ctypes.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()),
           "qdb_rk4_steps_c128")
    return y


def expm_steps(n, ops_rm, stat_rm, coeff, mu, times_mid_host: np.ndarray, squarings_host: np.ndarray, h, y, S,
               workspace=None):
    K = 0 if ops_rm is None else ops_rm.shape[0]
    B = y.shape[1]
    times_mid_host = np.ascontiguousarray(times_mid_host, dtype=np.float64)
    squarings_host = np.ascontiguousarray(squarings_host, dtype=np.int32)
    if workspace is None or workspace.numel() < workspace_bytes(WS_EXPM, n, K, B):
        workspace = torch.empty(workspace_bytes(WS_EXPM, n, K, B, S), dtype=torch.uint8, device=y.device)
    _check(lib().qdb_expm_steps_c128(n, K, B, S, _ptr(ops_rm, C, "ops_rm"), _ptr(stat_rm, C, "stat_rm"),
                                     _ptr(coeff, F, "coeff"), _ptr(mu, F, "mu"),
                                     times_mid_host.ctypes.data_as(ctypes.c_void_p),
                                     squarings_host.ctypes.data_as(ctypes.c_void_p), float(h), _ptr(y, C, "y"), B,
                                     ctypes.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()),
           "qdb_expm_steps_c128")
    return y


def magnus_steps(n, ops_rm, stat_rm, coeff, mu, times_host: np.ndarray, squarings_host: np.ndarray, h, y, S, magnus_order,
                 workspace=None):
    """S exponential steps at Magnus order 1, 2 or 3; times_host (S, order) node times, coeff (S, order, K)."""
    K = 0 if ops_rm is None else ops_rm.shape[0]
    B = y.shape[1]
    times_host = np.ascontiguousarray(times_host, dtype=np.float64)
    squarings_host = np.ascontiguousarray(squarings_host, dtype=np.int32)
    if times_host.size != S * magnus_order or squarings_host.size != S:
        raise QdbError(f"magnus_steps: {times_host.size} node times / {squarings_host.size} squarings for S={S}, "
                       f"order {magnus_order}")
    # a caller's workspace is used as it is when it holds at least one step (the library then works in smaller chunks)
    if workspace is None or workspace.numel() < workspace_bytes(WS_MAGNUS, n, K, B, 1) + 3 * S * 8 + 256:
        workspace = torch.empty(workspace_bytes(WS_MAGNUS, n, K, B, S), dtype=torch.uint8, device=y.device)
    _check(lib().qdb_magnus_steps_c128(n, K, B, S, int(magnus_order), _ptr(ops_rm, C, "ops_rm"), _ptr(stat_rm, C, "stat_rm"),
                                       _ptr(coeff, F, "coeff"), _ptr(mu, F, "mu"),
                                       times_host.ctypes.data_as(ctypes.c_void_p),
                                       squarings_host.ctypes.data_as(ctypes.c_void_p), float(h), _ptr(y, C, "y"), B,
                                       ctypes.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()),
           "qdb_magnus_steps_c128")
    return y


def step_propagators(n, ops_rm, stat_rm, coeff, mu, times_host: np.ndarray, squarings_host, h, S, kind, max_ws_bytes=1 << 31,
                     out=None):
    """Product of the S one-step propagators of an interval (kind 0: RK4, 1..3: expm at that Magnus order)."""
    K = 0 if ops_rm is None else ops_rm.shape[0]
    dev = (ops_rm if ops_rm is not None else stat_rm).device
    Q = 3 if kind == 0 else kind
    times_host = np.ascontiguousarray(times_host, dtype=np.float64)
    if times_host.size != S * Q:
        raise QdbError(f"step_propagators: {times_host.size} node times for S={S} steps of {Q} nodes")
    sq_ptr = None
    if kind != 0:
        squarings_host = np.ascontiguousarray(squarings_host, dtype=np.int32)
        if squarings_host.size != S:
            raise QdbError(f"step_propagators: {squarings_host.size} squarings for S={S}")
        sq_ptr = squarings_host.ctypes.data_as(ctypes.c_void_p)
    need = min(workspace_bytes(WS_PROP, n, K, 0, S), max(max_ws_bytes, workspace_bytes(WS_PROP, n, K, 0, 1)))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((n, n), dtype=C, device=dev)
    _check(lib().qdb_step_propagators_c128(n, K, S, int(kind), _ptr(ops_rm, C, "ops_rm"), _ptr(stat_rm, C, "stat_rm"),
                                           _ptr(coeff, F, "coeff"), _ptr(mu, F, "mu"),
                                           times_host.ctypes.data_as(ctypes.c_void_p), sq_ptr, float(h), _ptr(out, C, "P_total"),
                                           ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream()), "qdb_step_propagators_c128")
    return out


def magnus_terms(g: torch.Tensor, h: float, magnus_order: int, scale: float = 1.0):
    """scale * Omega(h) from the generators at the nodes, g of shape (order, n, n)."""
    n = g.shape[-1]
    out = torch.empty((n, n), dtype=C, device=g.device)
    ws = torch.empty(max(1, 7 * n * n * 16), dtype=torch.uint8, device=g.device)
    _check(lib().qdb_magnus_terms_c128(n, int(magnus_order), _ptr(g, C, "g"), float(h), float(scale), _ptr(out, C, "out"),
                                       ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream()), "qdb_magnus_terms_c128")
    return out


def expm(A: torch.Tensor, squarings: int, out=None):
    n = A.shape[0]
    if out is None:
        out = torch.empty_like(A)
    ws = torch.empty(6 * n * n * 16, dtype=torch.uint8, device=A.device)
    _check(lib().qdb_expm_c128(n, _ptr(A, C, "A"), int(squarings), _ptr(out, C, "out"), ctypes.c_void_p(ws.data_ptr()),
                               ws.numel(), _stream()), "qdb_expm_c128")
    return out


def rk4_table_layout(n: int, B: int) -> int:
    """Generator-table layout the fused shared-signal solve uses for this shape (PACKED or PACKED3M)."""
    return int(lib().qdb_rk4_table_layout(n, B))


def to_packed3m(table: torch.Tensor) -> torch.Tensor:
    """(T, npad*kpad) PACKED complex table -> (T, 3/2 npad*kpad) PACKED3M (complex plane + re+im plane)."""
    sums = (table.real + table.imag).contiguous()
    return torch.cat([table, torch.view_as_complex(sums.view(table.shape[0], -1, 2))], dim=1).contiguous()


def rk4_table_steps(n, table, h, y, S, layout=LAYOUT_PACKED):
    """The on-chip RK4 kernel alone, from a prebuilt generator table of 2S+1 entries in `layout`
    (LAYOUT_PACKED: (T, npad*kpad) -> 4-product kernels; LAYOUT_PACKED3M: (T, 3/2 npad*kpad) -> 3-product kernel)."""
    B = y.shape[1]
    if table.shape[0] < 2 * S + 1:
        raise QdbError(f"rk4_table_steps: table has {table.shape[0]} entries, need {2 * S + 1}")
    want = packed_elems(n) * (3 if layout == LAYOUT_PACKED3M else 2) // 2
    if table.shape[1] != want:
        raise QdbError(f"rk4_table_steps: table entries have {table.shape[1]} elements, layout {layout} needs {want}")
    _check(lib().qdb_rk4_table_steps_c128(n, B, S, _ptr(table, C, "table"), int(layout), float(h), _ptr(y, C, "y"), B,
                                          _stream()), "qdb_rk4_table_steps_c128")
    return y


def rk4_int8_preferred(n, B) -> bool:
    """True when a shared-signal RK4 solve of this shape runs on the int8 tensor-core emulation (rk4_ozaki_kernel)."""
    return bool(lib().qdb_rk4_int8_preferred(int(n), int(B)))


def rk4_ozaki_slice(n, table, layout=LAYOUT_ROWMAJOR, workspace=None):
    """Generator table (T, entry) -> the int8 slice planes rk4_ozaki_steps(..., table=None) consumes."""
    T = table.shape[0]
    need = int(lib().qdb_rk4_ozaki_workspace_bytes(max(1, (T - 1) // 2 + (T - 1) % 2)))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=table.device)
    _check(lib().qdb_rk4_ozaki_slice_c128(n, T, _ptr(table, C, "table"), int(layout), ctypes.c_void_p(workspace.data_ptr()),
                                          workspace.numel(), _stream()), "qdb_rk4_ozaki_slice_c128")
    return workspace


def rk4_ozaki_steps(n, table_rowmajor, h, y, S, workspace=None):
    """S RK4 steps with the contraction emulated on the int8 tensor cores, from a ROW-MAJOR generator table (2S+1, n*n) --
    or, with ``table_rowmajor=None``, from a ``workspace`` rk4_ozaki_slice filled with exactly 2S+1 entries."""
    B = y.shape[1]
    if table_rowmajor is not None and (table_rowmajor.shape[0] != 2 * S + 1 or table_rowmajor.shape[1] != n * n):
        raise QdbError(f"rk4_ozaki_steps: table shape {tuple(table_rowmajor.shape)}, need ({2 * S + 1}, {n * n})")
    need = int(lib().qdb_rk4_ozaki_workspace_bytes(S))
    if workspace is None or workspace.numel() < need:
        if table_rowmajor is None:
            raise QdbError("rk4_ozaki_steps: no table and no sliced workspace")
        workspace = torch.empty(need, dtype=torch.uint8, device=y.device)
    _check(lib().qdb_rk4_ozaki_steps_c128(n, B, S, _ptr(table_rowmajor, C, "table"), float(h), _ptr(y, C, "y"), B,
                                          ctypes.c_void_p(workspace.data_ptr()), workspace.numel(), _stream()),
           "qdb_rk4_ozaki_steps_c128")
    return y


def signal_table(K, terms: dict, samples, times, B=0, col_stride=0, scale=None, out=None, params_per_col=False):
    """``times``: float64 device tensor (T,), or a Python float (one time, no copy)."""
    """Coefficient table (T, K) (B == 0) or (T, K, B) of K channels from flattened term arrays on the device.

    ``terms``: dict of device tensors ``chan`` (int32), ``samp_off`` (int64), ``samp_len`` (int32), ``dt``, ``t0``,
    ``freq``, ``phase`` (float64), each of length nterms; ``samples``: complex128 sample storage."""
    scalar = not isinstance(times, torch.Tensor)
    T = 1 if scalar else int(times.shape[0])
    nterms = int(terms["chan"].shape[0])
    if out is None:
        out = torch.empty((T, K) if B == 0 else (T, K, B), dtype=F, device=samples.device)
    I32, I64 = torch.int32, torch.int64
    _check(lib().qdb_signal_table_f64(T, K, B, nterms, _ptr(terms["chan"], I32, "chan"), _ptr(terms["samp_off"], I64, "samp_off"),
                                      _ptr(terms["samp_len"], I32, "samp_len"), _ptr(terms["dt"], F, "dt"),
                                      _ptr(terms["t0"], F, "t0"), _ptr(terms["freq"], F, "freq"),
                                      _ptr(terms["phase"], F, "phase"), int(bool(params_per_col)), _ptr(samples, C, "samples"),
                                      int(col_stride),
                                      _ptr(scale, C, "scale"), None if scalar else _ptr(times, F, "times"),
                                      float(times) if scalar else 0.0, _ptr(out, F, "out"), _stream()),
           "qdb_signal_table_f64")
    return out


def outcome_probabilities(y, outcome_of, n_out: int, normalize: bool = True, out=None):
    """(n_out, B) outcome probabilities of the columns of y (n, B); outcome_of: int32 (n,) bin of every basis state."""
    n, B = y.shape
    if out is None:
        out = torch.empty((n_out, B), dtype=F, device=y.device)
    _check(lib().qdb_outcome_probabilities_f64(n, B, int(n_out), _ptr(y, C, "y"), B, _ptr(outcome_of, torch.int32, "outcome_of"),
                                               int(bool(normalize)), _ptr(out, F, "out"), _stream()),
           "qdb_outcome_probabilities_f64")
    return out


def lindblad_supported(n: int) -> bool:
    """True when the on-chip non-vectorised Lindblad kernels take this dimension (n <= 32)."""
    return bool(lib().qdb_lindblad_supported(int(n)))


def lindblad_rhs(n, m1_packed, m2t_packed, diss_packed, gamma, mu, t, rho, out=None):
    """One Lindblad RHS evaluation on rho (B, n, n); m1 / m2t: packed M1(t), M2(t)^T; diss_packed (J, npad*kpad) or None."""
    B = rho.shape[0]
    J = 0 if diss_packed is None else diss_packed.shape[0]
    if out is None:
        out = torch.empty_like(rho)
    _check(lib().qdb_lindblad_rhs_c128(n, J, B, _ptr(m1_packed, C, "m1"), _ptr(m2t_packed, C, "m2t"), _ptr(diss_packed, C, "diss"),
                                       _ptr(gamma, F, "gamma"), _ptr(mu, F, "mu"), float(t), _ptr(rho, C, "rho_in"),
                                       _ptr(out, C, "rho_out"), _stream()), "qdb_lindblad_rhs_c128")
    return out


def lindblad_rk4_steps(n, m1_table, m2t_table, diss_packed, gamma_table, mu, times_dev, h, rho, S):
    """S fused RK4 steps in place on rho (B, n, n); tables (2S+1, npad*kpad) packed, gamma_table (2S+1, J) or None."""
    B = rho.shape[0]
    J = 0 if diss_packed is None else diss_packed.shape[0]
    if m1_table.shape[0] < 2 * S + 1 or m2t_table.shape[0] < 2 * S + 1:
        raise QdbError(f"lindblad_rk4_steps: tables have {m1_table.shape[0]} / {m2t_table.shape[0]} entries, need {2 * S + 1}")
    _check(lib().qdb_lindblad_rk4_steps_c128(n, J, B, S, _ptr(m1_table, C, "m1_table"), _ptr(m2t_table, C, "m2t_table"),
                                             _ptr(diss_packed, C, "diss"), _ptr(gamma_table, F, "gamma_table"), _ptr(mu, F, "mu"),
                                             _ptr(times_dev, F, "times"), float(h), _ptr(rho, C, "rho"), _stream()),
           "qdb_lindblad_rk4_steps_c128")
    return rho


def dmma_probe(iters: int = 20000, reps: int = 5) -> float:
    """Measured fp64 tensor-pipe peak in TFLOP/s (best of `reps`, CUDA events)."""
    sink = torch.zeros(1, dtype=F, device="cuda")
    flops = ctypes.c_double(0.0)
    best = float("inf")
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _check(lib().qdb_dmma_probe(_ptr(sink, F, "sink"), int(iters), ctypes.byref(flops), _stream()), "qdb_dmma_probe")
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            best = min(best, e0.elapsed_time(e1))
    return flops.value / best * 1e-9
