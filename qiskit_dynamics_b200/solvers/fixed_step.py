"""Fixed-step drivers: the step grid of the reference (host, NumPy) and the fused device loops.

* ``merge_t_args`` / ``trim_t_results`` / ``get_fixed_step_sizes`` keep the reference's exact
  semantics (solvers/solver_utils.py:46-119, solvers/fixed_step_solvers.py:616-653).
* ``rk4_model_solve`` / ``expm_model_solve`` own the step loop for model generators: per
  integration interval they build the stage-time grid with the reference's accumulation, evaluate
  the SignalList once into a (T, K) table, and make ONE C-ABI call (qdb_rk4_steps_c128 /
  qdb_expm_steps_c128) -- replacing the Python hot loop at solvers/fixed_step_solvers.py:448-454.
* ``RK4_solver`` / ``scipy_expm_solver`` keep the reference's generic callable protocol
  (``rhs(t, y)`` / ``generator(t)`` as arbitrary Python functions returning device tensors); this
  host-driven route cannot be fused and is not on the measured path.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from scipy.integrate._ivp.ivp import OdeResult

from .. import _abi
from ..arrays import asarray, asreal, stage_to_device
from ..exceptions import QiskitError

EXPM_THETA = 0.7  # 1-norm radius of the degree-16 Taylor polynomial (csrc/expm.cu)


# ---------------------------------------------------------------------------------------------
# step grid (host)
# ---------------------------------------------------------------------------------------------


def merge_t_args(t_span, t_eval=None) -> np.ndarray:
    """[t_span[0], *t_eval, t_span[1]] after validation; t_span itself when t_eval is None."""
    if t_eval is None:
        return t_span
    t_span = np.array(t_span)
    t_eval = np.array(t_eval)
    lo, hi = np.min(t_span), np.max(t_span)
    direction = np.sign(t_span[1] - t_span[0])
    if t_eval.ndim > 1:
        raise ValueError("t_eval must be 1 dimensional.")
    if np.min(t_eval) < lo or np.max(t_eval) > hi:
        raise ValueError("t_eval entries must lie in t_span.")
    if np.any(direction * np.diff(t_eval) < 0.0):
        raise ValueError("t_eval must be ordered according to the direction of integration.")
    return np.append(np.append(t_span[0], t_eval), t_span[1])


def trim_t_results(results: OdeResult, t_eval=None) -> OdeResult:
    """Drop the two end points that merge_t_args added."""
    if t_eval is None:
        return results
    results.t = results.t[1:-1]
    results.y = results.y[1:-1]
    return results


def get_fixed_step_sizes(t_span, t_eval, max_dt) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(t_list, h_list, n_steps_list): fewest equal steps per interval with |h| <= max_dt."""
    t_list = np.array(merge_t_args(np.array(t_span), t_eval))
    max_dt = np.array(max_dt)
    deltas = np.diff(t_list)
    n_steps = np.abs(deltas / max_dt).astype(int)
    for i, (delta, n) in enumerate(zip(deltas, n_steps)):
        if n == 0:
            n_steps[i] = 1
        elif np.abs(delta / n) / max_dt > 1 + 1e-15:  # guard against the truncation above
            n_steps[i] = n + 1
    return t_list, np.array(deltas / n_steps), n_steps


def stage_time_grid(t0: float, h: float, n_steps: int) -> np.ndarray:
    """[t_0, t_0 + h/2, t_1, t_1 + h/2, ..., t_S] with t_{i+1} = t_i + h accumulated sequentially
    (cumsum is a sequential add), exactly the floats the reference's loop produces
    (fixed_step_solvers.py:64-66, 453) -- this matters for DiscreteSignal bin edges."""
    starts = np.cumsum(np.concatenate([[float(t0)], np.full(int(n_steps), float(h))]))
    grid = np.empty(2 * int(n_steps) + 1)
    grid[0::2] = starts
    grid[1::2] = starts[:-1] + 0.5 * h
    return grid


# ---------------------------------------------------------------------------------------------
# fused model solves
# ---------------------------------------------------------------------------------------------


def _columns(y0: torch.Tensor):
    if y0.ndim == 1:
        return y0.reshape(-1, 1).contiguous(), y0.shape
    if y0.ndim == 2:
        return y0.contiguous(), y0.shape
    raise QiskitError("y0 must be a vector or a matrix whose columns are states.")


def rk4_model_solve(model, t_span, y0_fb: torch.Tensor, max_dt, t_eval=None,
                    column_coefficients: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                    workspace_bytes: int = 1 << 30, sweep_table_bytes: int = 1 << 30) -> OdeResult:
    """RK4 on a model generator, state already in the frame basis.

    ``column_coefficients(times) -> (T, K, B)`` (NumPy array or device tensor) switches to sweep mode
    (per-column signal values).
    """
    coll = model._collection()
    n = coll.dim
    y, shape = _columns(y0_fb)
    if y.shape[0] != n:
        raise QiskitError(f"y0 has leading dimension {y.shape[0]}, model dimension is {n}.")
    y = y.clone()
    B = y.shape[1]
    mu = model._frame_freqs()
    t_list, h_list, n_list = get_fixed_step_sizes(t_span, t_eval, max_dt)
    ops, stat = coll.operators, coll.static_operator
    use_packed = _abi.npad(n) <= 256
    ops_p, stat_p = coll.packed() if use_packed else (None, None)
    K = coll.num_operators
    ws = None
    ys = [y.reshape(shape).clone()]
    for t0, h, S in zip(t_list, h_list, n_list):
        S = int(S)
        times = stage_time_grid(t0, h, S)
        if column_coefficients is not None:
            # sweep mode: the (T, K, B) table is built (host or device) and consumed in chunks of steps so that it
            # never exceeds sweep_table_bytes of HBM, whatever the number of steps
            per_step = 2 * K * B * 8
            chunk = max(1, min(S, int(sweep_table_bytes // max(per_step, 1))))
            for s0 in range(0, S, chunk):
                Sc = min(chunk, S - s0)
                tchunk = times[2 * s0: 2 * (s0 + Sc) + 1]
                table = column_coefficients(tchunk)
                if isinstance(table, torch.Tensor):
                    coeff = table.to(device=y.device, dtype=torch.float64).contiguous()
                else:
                    coeff = asreal(np.ascontiguousarray(table, dtype=np.float64), y.device)
                if tuple(coeff.shape) != (tchunk.shape[0], K, B):
                    raise QiskitError(f"per-column signal table has shape {tuple(coeff.shape)}, expected {(tchunk.shape[0], K, B)}")
                need = _abi.workspace_bytes(_abi.WS_RK4, n, K, B, Sc)
                if ws is None or ws.numel() < need:
                    ws = torch.empty(need, dtype=torch.uint8, device=y.device)
                _abi.rk4_steps(n, ops, stat, ops_p, stat_p, coeff, mu, tchunk, float(h), y, Sc, per_col=True, workspace=ws)
        else:
            coeff = model._signal_table_device(times, y.device) if hasattr(model, "_signal_table_device") else None
            if coeff is None:  # arbitrary Python envelopes: one vectorised host evaluation
                table = model._signal_table(times)
                coeff = None if table is None else asreal(table, y.device)
            need = _abi.workspace_bytes(_abi.WS_RK4, n, K, B, S)
            need = min(need, max(workspace_bytes, _abi.workspace_bytes(_abi.WS_RK4, n, K, B, 1)))
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, dtype=torch.uint8, device=y.device)
            _abi.rk4_steps(n, ops, stat, ops_p, stat_p, coeff, mu, times, float(h), y, S, per_col=False, workspace=ws)
        ys.append(y.reshape(shape).clone())
    return trim_t_results(OdeResult(t=np.array(t_list), y=torch.stack(ys)), t_eval)


def lindblad_rk4_model_solve(model, t_span, y0_fb: torch.Tensor, max_dt, t_eval=None, table_bytes: int = 1 << 28) -> OdeResult:
    """RK4 on a NON-vectorised LindbladModel (dim <= 32), density matrices already in the frame basis: y0 (n, n) or a
    batch (l, n, n).  Per integration interval (chunked so that the two generator tables stay below ``table_bytes``):
    the signal values on the stage-time grid go up once, two qdb_generator_c128 launches build M1(t) = A + B and
    M2(t)^T = (A - B)^T for every stage time, and ONE qdb_lindblad_rk4_steps_c128 launch runs the steps with every density
    matrix resident on chip -- replacing the host loop of RK4_solver over LindbladModel.evaluate_rhs
    (solvers/fixed_step_solvers.py:43-77, 441-454; models/lindblad_model.py:477-538)."""
    coll = model._operator_collection
    n = model.dim
    single = y0_fb.ndim == 2
    rho = (y0_fb.unsqueeze(0) if single else y0_fb).contiguous().clone()
    if rho.ndim != 3 or tuple(rho.shape[-2:]) != (n, n):
        raise QiskitError(f"y0 must be (n, n) or (l, n, n) with n = {n}.")
    f = coll.fused_operands()
    mu = model.rotating_frame.frame_freqs
    t_list, h_list, n_list = get_fixed_step_sizes(t_span, t_eval, max_dt)
    per_step = 2 * 2 * _abi.packed_elems(n) * 16
    ys = [rho[0].clone() if single else rho.clone()]
    for t0, h, S in zip(t_list, h_list, n_list):
        S = int(S)
        times = stage_time_grid(t0, h, S)
        ham_tab, dis_tab = model._signal_table_parts(times)
        chunk = max(1, min(S, int(table_bytes // per_step)))
        for s0 in range(0, S, chunk):
            Sc = min(chunk, S - s0)
            sl = slice(2 * s0, 2 * (s0 + Sc) + 1)
            m1, m2t, gamma = coll.fused_tables(None if ham_tab is None else ham_tab[sl], None if dis_tab is None else dis_tab[sl],
                                               rho.device)
            if m1.shape[0] == 1:  # no time-dependent operator at all: one entry serves every stage
                m1, m2t = m1.expand(2 * Sc + 1, -1).contiguous(), m2t.expand(2 * Sc + 1, -1).contiguous()
            times_dev = None if mu is None else stage_to_device(times[sl], rho.device)
            _abi.lindblad_rk4_steps(n, m1, m2t, f["diss"], gamma, mu, times_dev, float(h), rho, Sc)
        ys.append(rho[0].clone() if single else rho.clone())
    return trim_t_results(OdeResult(t=np.array(t_list), y=torch.stack(ys)), t_eval)


def magnus_nodes(magnus_order: int) -> np.ndarray:
    """Generator evaluation points of one exponential step, as fractions of h: the midpoint (order 1) or the
    2- / 3-point Gauss-Legendre nodes (fixed_step_solvers.py:346, 350-351, 367-369)."""
    if magnus_order == 1:
        return np.array([0.5])
    if magnus_order == 2:
        return np.array([0.5 - np.sqrt(3) / 6, 0.5 + np.sqrt(3) / 6])
    if magnus_order == 3:
        return np.array([0.5 - np.sqrt(15) / 10, 0.5, 0.5 + np.sqrt(15) / 10])
    raise QiskitError("Only magnus_order 1, 2, and 3 are supported.")


def magnus_norm_bound(b: np.ndarray, h: float, magnus_order: int) -> np.ndarray:
    """Upper bound on ||Omega||_1 of a step from bounds ``b[..., q]`` on ||G(t_q)||_1 at its nodes
    (sub-multiplicativity: ||[X, Y]|| <= 2 ||X|| ||Y||)."""
    h = abs(float(h))
    if magnus_order == 1:
        return h * b[..., 0]
    if magnus_order == 2:
        return h * (b[..., 0] + b[..., 1]) / 2 + (np.sqrt(3) / 12) * h * h * 2 * b[..., 0] * b[..., 1]
    a1 = h * b[..., 1]
    a2 = (np.sqrt(15) / 3) * h * (b[..., 2] + b[..., 0])
    a3 = (10.0 / 3) * h * (b[..., 2] + 2 * b[..., 1] + b[..., 0])
    comm1 = 2 * a1 * a2
    comm2 = 2 * (2 * a3 + comm1) * a1 / 60
    return a1 + a3 / 12 + 2 * (20 * a1 + a3 + comm1) * (a2 + comm2) / 240


def expm_squarings(model, coeff_table: Optional[np.ndarray], h: float, magnus_order: int = 1) -> np.ndarray:
    """Squarings per step from the bound ||G(t)||_1 <= ||G_d||_1 + sum |c_j| ||G_j||_1 at every node (the frame
    phases have modulus 1) pushed through the Magnus exponent, so that no device->host norm read is needed
    inside the loop.  ``coeff_table``: (S * order, K) signal values at the node times, or None."""
    s_norm, o_norms = model._collection().norms1()
    rows = magnus_order if coeff_table is None else coeff_table.shape[0]
    b = np.full(rows, s_norm, dtype=float)
    if coeff_table is not None and o_norms.size:
        b = b + np.abs(coeff_table) @ o_norms
    bound = magnus_norm_bound(b.reshape(-1, magnus_order), h, magnus_order)
    with np.errstate(divide="ignore"):
        sq = np.ceil(np.log2(np.maximum(bound, 1e-300) / EXPM_THETA))
    return np.maximum(sq, 0).astype(np.int32)


def expm_model_solve(model, t_span, y0_fb: torch.Tensor, max_dt, t_eval=None, magnus_order: int = 1) -> OdeResult:
    """y <- expm(Omega) y per step on a model generator, Omega the Magnus exponent of order 1, 2 or 3 built from
    the generator at the step's nodes (order 1: h G(t + h/2))."""
    nodes = magnus_nodes(magnus_order)
    coll = model._collection()
    n = coll.dim
    y, shape = _columns(y0_fb)
    if y.shape[0] != n:
        raise QiskitError(f"y0 has leading dimension {y.shape[0]}, model dimension is {n}.")
    y = y.clone()
    mu = model._frame_freqs()
    t_list, h_list, n_list = get_fixed_step_sizes(t_span, t_eval, max_dt)
    ws = None
    ys = [y.reshape(shape).clone()]
    for t0, h, S in zip(t_list, h_list, n_list):
        S = int(S)
        starts = stage_time_grid(t0, h, S)[0::2][:-1]
        if magnus_order == 1:
            times = (starts + (h / 2)).reshape(S, 1)
        else:
            times = starts[:, None] + nodes[None, :] * h  # t0 + c_q * h, the reference's expression
        table = model._signal_table(times.reshape(-1))
        coeff = None if table is None else asreal(table, y.device)
        sq = expm_squarings(model, table, float(h), magnus_order)
        if table is None:
            sq = np.repeat(sq, S)
        if magnus_order == 1:
            need = _abi.workspace_bytes(_abi.WS_EXPM, n, coll.num_operators, y.shape[1], S)  # room for batched propagators
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, dtype=torch.uint8, device=y.device)
            _abi.expm_steps(n, coll.operators, coll.static_operator, coeff, mu, times.reshape(-1), sq, float(h), y, S, workspace=ws)
        else:
            need = _abi.workspace_bytes(_abi.WS_MAGNUS, n, coll.num_operators, y.shape[1], S)
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, dtype=torch.uint8, device=y.device)
            _abi.magnus_steps(n, coll.operators, coll.static_operator, coeff, mu, times, sq, float(h), y, S, magnus_order,
                              workspace=ws)
        ys.append(y.reshape(shape).clone())
    return trim_t_results(OdeResult(t=np.array(t_list), y=torch.stack(ys)), t_eval)


def parallel_model_solve(model, t_span, y0_fb: torch.Tensor, max_dt, t_eval=None, kind: str = "RK4", magnus_order: int = 1,
                         workspace_bytes: int = 1 << 31) -> OdeResult:
    """Time-parallel solve of a model generator (the reference's jax_RK4_parallel / jax_expm_parallel,
    fixed_step_solvers.py:206-244, 279-311, 524-613): per integration interval ONE C-ABI call builds the propagators
    of all its steps side by side and multiplies them (qdb_step_propagators_c128), one GEMM applies the product to the
    state batch.  Step times are ``t + h * arange(n_steps)`` as in the reference's template."""
    if kind not in ("RK4", "expm"):
        raise QiskitError(f"unknown time-parallel stepper {kind}")
    order = 0 if kind == "RK4" else int(magnus_order)
    nodes = np.array([0.0, 0.5, 1.0]) if kind == "RK4" else magnus_nodes(order)
    coll = model._collection()
    n = coll.dim
    y, shape = _columns(y0_fb)
    if y.shape[0] != n:
        raise QiskitError(f"y0 has leading dimension {y.shape[0]}, model dimension is {n}.")
    mu = model._frame_freqs()
    t_list, h_list, n_list = get_fixed_step_sizes(t_span, t_eval, max_dt)
    ys = [y.reshape(shape).clone()]
    for t0, h, S in zip(t_list, h_list, n_list):
        S = int(S)
        starts = t0 + h * np.arange(S)
        if kind == "RK4":  # t, t + 0.5 h, t + h -- the expressions of the reference's take_step
            times = np.stack([starts, starts + 0.5 * h, starts + h], axis=1)
        elif order == 1:
            times = (starts + (h / 2)).reshape(S, 1)
        else:
            times = starts[:, None] + nodes[None, :] * h
        table = model._signal_table(times.reshape(-1))
        coeff = None if table is None else asreal(table, y.device)
        sq = None
        if kind == "expm":
            sq = expm_squarings(model, table, float(h), order)
            if table is None:
                sq = np.repeat(sq, S)
        P = _abi.step_propagators(n, coll.operators, coll.static_operator, coeff, mu, times, sq, float(h), S, order,
                                  max_ws_bytes=workspace_bytes)
        y = _abi.zgemm(P, y)
        ys.append(y.reshape(shape).clone())
    return trim_t_results(OdeResult(t=np.array(t_list), y=torch.stack(ys)), t_eval)


# ---------------------------------------------------------------------------------------------
# generic callable protocol (host-driven; not fused)
# ---------------------------------------------------------------------------------------------


def fixed_step_solver_template(take_step: Callable, rhs_func: Callable, t_span, y0, max_dt, t_eval=None) -> OdeResult:
    y0 = asarray(y0)
    t_list, h_list, n_list = get_fixed_step_sizes(t_span, t_eval, max_dt)
    ys = [y0]
    for t0, h, n in zip(t_list, h_list, n_list):
        y = ys[-1]
        t = t0
        for _ in range(int(n)):
            y = take_step(rhs_func, t, y, h)
            t = t + h
        ys.append(y)
    return trim_t_results(OdeResult(t=np.array(t_list), y=torch.stack(ys)), t_eval)


def RK4_solver(rhs: Callable, t_span, y0, max_dt, t_eval=None) -> OdeResult:
    """Classical RK4 for an arbitrary ``rhs(t, y)`` returning device tensors."""
    div6 = 1.0 / 6

    def take_step(f, t, y, h):
        h2 = 0.5 * h
        k1 = asarray(f(t, y))
        k2 = asarray(f(t + h2, y + h2 * k1))
        k3 = asarray(f(t + h2, y + h2 * k2))
        k4 = asarray(f(t + h, y + h * k3))
        return y + div6 * h * (k1 + 2 * k2 + 2 * k3 + k4)

    return fixed_step_solver_template(take_step, rhs, t_span, y0, max_dt, t_eval)


def scipy_expm_solver(generator: Callable, t_span, y0, max_dt, t_eval=None, magnus_order: int = 1) -> OdeResult:
    """Exponential stepper (Magnus order 1, 2 or 3) for an arbitrary ``generator(t)`` returning an (n, n) device
    tensor; the Magnus exponent (qdb_magnus_terms_c128) and the exponential itself (qdb_expm_c128, Taylor
    scaling-and-squaring) are built on the device."""
    nodes = magnus_nodes(magnus_order)

    def take_step(gen, t, y, h):
        if magnus_order == 1:
            A = (asarray(gen(t + (h / 2))) * h).contiguous()
        else:
            g = torch.stack([asarray(gen(t + c * h)) for c in nodes]).contiguous()
            A = _abi.magnus_terms(g, float(h), magnus_order)
        norm = float(torch.linalg.matrix_norm(A, 1))
        sq = max(0, int(np.ceil(np.log2(max(norm, 1e-300) / EXPM_THETA))))
        P = _abi.expm(A, sq)
        if y.ndim == 1:
            return _abi.zgemm(P, y.reshape(-1, 1).contiguous()).reshape(-1)
        return _abi.zgemm(P, y.contiguous())

    return fixed_step_solver_template(take_step, generator, t_span, y0, max_dt, t_eval)
