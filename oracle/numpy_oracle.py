"""CPU oracle for the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-NumPy restatement of the reference's algorithm for the path named by
BASELINE.json `north_star`.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; the product package `qiskit_dynamics_b200` never does
(it talks to the CUDA C-ABI library only and fails loudly without it).

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function below against
the fixtures in `tests/golden/*.npz`, which were produced by running the unmodified
reference (qiskit-dynamics 0.6.0 @ 6b54df2f, NumPy path) through `oracle/ref_shim.py` with
`tests/golden/make_golden.py`; when `/root/reference` is mounted the same test module also
compares this oracle against the live reference.  One exception: the time-parallel template
(`parallel_solve`, `rk4_step_propagator`) restates code the reference runs on JAX only, which is
absent here, so no fixture exists for it; it is pinned INDIRECTLY -- the same step propagators in a
different association order must reproduce the fixture-pinned sequential solvers
(`test_parallel_template_agrees_with_sequential`, 1e-13).

Third-party arithmetic at the boundary: NumPy/OpenBLAS (`tensordot`, `matmul`, ufuncs) and
`scipy.linalg.expm` -- exactly the calls the reference makes, so that this port is also a
faithful *timing* stand-in for the reference's NumPy path (same BLAS calls, same shapes).

Each function cites the reference file:line it follows (paths relative to
/root/reference/qiskit_dynamics/).
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import expm as _scipy_expm

# ----------------------------------------------------------------------------------------------
# a6  signals
# ----------------------------------------------------------------------------------------------


class SigSpec:
    """One elementary signal  s(t) = Re[f(t) exp(i(2 pi nu t + phi))].

    `envelope` is a number, a vectorised callable, or the tuple
    ("discrete", dt, samples, start_time) for a piecewise-constant envelope.
    Follows signals/signals.py:70-155 (Signal) and :268-313 (DiscreteSignal).
    """

    def __init__(self, envelope, carrier_freq=0.0, phase=0.0):
        self.envelope = envelope
        self.carrier_freq = float(carrier_freq)
        self.phase = float(phase)


def discrete_envelope(dt: float, samples: np.ndarray, start_time: float, t) -> np.ndarray:
    """Piecewise-constant envelope lookup; signals/signals.py:297-311.

    Uses NumPy *float floor division* on purpose (bin-edge semantics): the index
    is clip(int((t - t0) // dt), -1, N) into samples padded with one trailing zero, so both
    out-of-range sides read zero.
    """
    samples = np.asarray(samples)
    padded = np.append(samples, np.zeros((1,) + samples.shape[1:], dtype=samples.dtype), axis=0)
    t = np.asarray(t)
    idx = np.clip(np.array((t - start_time) // dt, dtype=int), -1, len(samples))
    return padded[idx]


def _envelope_values(spec: SigSpec, t) -> np.ndarray:
    env = spec.envelope
    if isinstance(env, tuple) and len(env) == 4 and env[0] == "discrete":
        _, dt, samples, start = env
        return discrete_envelope(dt, samples, start, t)
    if callable(env):
        return env(t)
    return np.asarray(env) * np.ones_like(t)  # signals/signals.py:102


def signal_complex_value(spec: SigSpec, t) -> np.ndarray:
    """f(t) exp(i 2 pi nu t + i phi); signals/signals.py:148-151 (arg built as in :126-141)."""
    carrier_arg = 1j * 2 * np.pi * np.asarray(spec.carrier_freq)
    phase_arg = 1j * np.asarray(spec.phase)
    return _envelope_values(spec, t) * np.exp(carrier_arg * t + phase_arg)


def signal_list_values(specs: Sequence[Sequence[SigSpec] | SigSpec], t) -> np.ndarray:
    """SignalList.__call__: real part of each entry's SignalSum; signals/signals.py:574-577,
    :787-803.  Scalar t -> (K,), array t of shape (T,) -> (T, K)."""
    t = np.asarray(t, dtype=float)
    cols = []
    for entry in specs:
        terms = entry if isinstance(entry, (list, tuple)) else [entry]
        # SignalSum.complex_value: sum over terms of envelope * exp(t*carrier_arg + phase_arg)
        carrier_arg = 1j * 2 * np.pi * np.array([s.carrier_freq for s in terms])
        phase_arg = 1j * np.array([s.phase for s in terms])
        env = np.moveaxis(np.asarray([_envelope_values(s, t) for s in terms]), 0, -1)
        exp_phases = np.exp(np.expand_dims(t, -1) * carrier_arg + phase_arg)
        cols.append(np.real(np.sum(env * exp_phases, axis=-1)))
    return np.moveaxis(np.asarray(cols), 0, -1)


# ----------------------------------------------------------------------------------------------
# a3 / A.2  rotating frame
# ----------------------------------------------------------------------------------------------


def frame_decompose(frame_operator, atol=1e-10, rtol=1e-10):
    """Returns (frame_diag d, U) with d = -i*eigvals, U = eigenvectors (None for 1-d input).

    models/rotating_frame.py:87-108 and _enforce_anti_herm :648-660: Hermitian input H is
    converted to F = -iH; anti-Hermitian input is used as is.
    """
    if frame_operator is None:
        return None, None
    F = np.asarray(frame_operator)
    if F.ndim == 1:
        if np.allclose(F, F.conj(), atol=atol, rtol=rtol):
            F = -1j * F
        elif not np.allclose(F, -F.conj(), atol=atol, rtol=rtol):
            raise ValueError("frame_operator must be Hermitian or anti-Hermitian")
        return F, None
    if np.allclose(F, F.conj().T, atol=atol, rtol=rtol):
        F = -1j * F
    elif not np.allclose(1j * F, (1j * F).conj().T, atol=atol, rtol=rtol):
        raise ValueError("frame_operator must be Hermitian or anti-Hermitian")
    lam, U = np.linalg.eigh(1j * F)
    return -1j * lam, U


def state_into_frame(d, t, y):
    """Row i of y times exp(-d_i t), in the frame basis; models/rotating_frame.py:255."""
    return (np.exp(d * (-t)) * y.transpose()).transpose()


def state_out_of_frame(d, t, y):
    """state_into_frame with time reversed; models/rotating_frame.py:284."""
    return state_into_frame(d, -t, y)


def operator_into_frame(d, t, op):
    """op (frame basis) Hadamard outer(conj(e), e), e = exp(d t); rotating_frame.py:350-353."""
    e = np.exp(d * t)
    return op * (e.conj().reshape(len(d), 1) * e)


def vectorized_map_into_frame(d, t, op):
    """(n^2,n^2) superoperator in the frame (frame basis in and out); rotating_frame.py:568-577."""
    n = len(d)
    e = np.exp(d * t)
    tau = (e.conj().reshape(n, 1) * e).flatten()
    return np.outer(tau.conj(), tau) * op


# ----------------------------------------------------------------------------------------------
# A.3  stored operators of a HamiltonianModel / GeneratorModel
# ----------------------------------------------------------------------------------------------


def generator_model_operators(static_operator, operators, frame_operator, hamiltonian=True):
    """Frame-basis operators exactly as the model stores them.

    hamiltonian=True folds the -i (models/hamiltonian_model.py:97-111).  Then
    G_j = U^dag G_j U (generator_model.py:363-365 -> rotating_frame.py:195) and
    G_d = U^dag G_d U - diag(d) (generator_model.py:336-340 -> rotating_frame.py:463-474 at t=0),
    or diag(-d) when only a frame is given (generator_model.py:329-334).
    Returns (G_d or None, G (K,n,n) or None, d or None, U or None).
    """
    d, U = frame_decompose(frame_operator)
    fac = -1j if hamiltonian else 1.0
    Gd = None if static_operator is None else fac * np.asarray(static_operator, dtype=complex)
    G = None if operators is None else fac * np.asarray(operators, dtype=complex)

    def into_basis(op):
        if U is None or op is None:
            return op
        return np.matmul(U.conj().T, np.matmul(op, U))

    G = into_basis(G)
    if Gd is None:
        if d is not None:
            Gd = np.diag(-d)
    else:
        Gd = into_basis(Gd)
        if d is not None:
            # _conjugate_and_add at t = 0: multiply by outer(conj(e), e) with e = exp(0) = 1
            e = np.exp(d * 0.0)
            Gd = Gd * (e.conj().reshape(len(d), 1) * e) + (-np.diag(d))
    return Gd, G, d, U


# ----------------------------------------------------------------------------------------------
# a1 / a2  operator collection
# ----------------------------------------------------------------------------------------------


def collection_evaluate(coeffs, ops, static):
    """G_d + sum_j c_j G_j via tensordot; models/operator_collections.py:101-122 and
    arraylias/register_functions/linear_combo.py:30-32."""
    if static is not None and ops is not None:
        return np.tensordot(coeffs, ops, axes=1) + static
    if ops is not None:
        return np.tensordot(coeffs, ops, axes=1)
    if static is not None:
        return static
    raise ValueError("collection with neither static operator nor operators")


def collection_evaluate_rhs(coeffs, ops, static, y):
    """(G_d + sum_j c_j G_j) @ y; models/operator_collections.py:124-134."""
    return np.matmul(collection_evaluate(coeffs, ops, static), y)


# ----------------------------------------------------------------------------------------------
# a4 / a5  model evaluation in the frame basis (what the solver loop calls)
# ----------------------------------------------------------------------------------------------


def model_rhs(t, y, specs, ops, static, d):
    """GeneratorModel.evaluate_rhs with in_frame_basis=True; models/generator_model.py:301-314."""
    c = None if specs is None else signal_list_values(specs, t)
    if d is None:
        return collection_evaluate_rhs(c, ops, static, y)
    out = state_out_of_frame(d, t, y)
    out = collection_evaluate_rhs(c, ops, static, out)
    return state_into_frame(d, t, out)


def model_generator(t, specs, ops, static, d):
    """GeneratorModel.evaluate with in_frame_basis=True; models/generator_model.py:274-279."""
    c = None if specs is None else signal_list_values(specs, t)
    G = collection_evaluate(c, ops, static)
    return G if d is None else operator_into_frame(d, t, G)


# ----------------------------------------------------------------------------------------------
# a10  vectorised Lindblad
# ----------------------------------------------------------------------------------------------


def vec_commutator(A):
    """-i (I kron A - A^T kron I), column stacking; models/model_utils.py:31-71."""
    A = np.asarray(A)
    iden = np.eye(A.shape[-1])
    At = np.swapaxes(A, -1, -2)
    return -1j * (np.kron(iden, A) - np.kron(At, iden))


def vec_dissipator(L):
    """conj(L) kron L - (I kron L^dag L + (L^dag L)^T kron I)/2; models/model_utils.py:74-118."""
    L = np.asarray(L)
    iden = np.eye(L.shape[-1])
    Lc = L.conj()
    LdL = np.swapaxes(Lc, -1, -2) @ L
    LdLt = np.swapaxes(LdL, -1, -2)
    return np.kron(Lc, iden) @ np.kron(iden, L) - 0.5 * (np.kron(iden, LdL) + np.kron(LdLt, iden))


def _batched(fn, arr):
    arr = np.asarray(arr)
    if arr.ndim == 2:
        return fn(arr)
    return np.asarray([fn(a) for a in arr])


def lindblad_model_operators(static_hamiltonian, hamiltonian_operators, static_dissipators,
                             dissipator_operators, frame_operator):
    """Frame-basis Lindblad operators as LindbladModel stores them; lindblad_model.py:173-204.

    The static Hamiltonian keeps no -i: it is multiplied by -i, the frame is subtracted, and it
    is multiplied back by +i (:173-181).  Returns (H_d, H_ops, D_static, D_ops, d, U).
    """
    d, U = frame_decompose(frame_operator)

    def into_basis(op):
        if op is None:
            return None
        op = np.asarray(op, dtype=complex)
        if U is None:
            return op
        return np.matmul(U.conj().T, np.matmul(op, U))

    Hd = None
    if static_hamiltonian is not None or d is not None:
        Gd, _, _, _ = generator_model_operators(static_hamiltonian, None, frame_operator, True)
        Hd = None if Gd is None else 1j * Gd
    return (Hd, into_basis(hamiltonian_operators), into_basis(static_dissipators),
            into_basis(dissipator_operators), d, U)


def vectorized_lindblad_collection(Hd, Hops, Dstatic, Dops):
    """Superoperator collection (static (n^2,n^2), operators (K_h+K_d,n^2,n^2));
    models/operator_collections.py:887-938."""
    static = None
    if Hd is not None:
        static = vec_commutator(Hd)
    if Dstatic is not None:
        s = np.sum(_batched(vec_dissipator, Dstatic), axis=0) if np.asarray(Dstatic).ndim == 3 \
            else vec_dissipator(Dstatic)
        static = s if static is None else static + s
    parts = []
    if Hops is not None:
        parts.append(_batched(vec_commutator, Hops))
    if Dops is not None:
        parts.append(_batched(vec_dissipator, Dops))
    ops = None if not parts else np.concatenate(parts, axis=0)
    return static, ops


def lindblad_rhs_matrix(ham_c, dis_c, Hd, Hops, Dstatic, Dops, rho):
    """Non-vectorised Lindblad RHS in (A+B)rho + rho(A-B) + sum L rho L^dag form;
    models/operator_collections.py:451-567.  rho is (n,n) or (B,n,n)."""
    H = None
    if Hd is not None or Hops is not None:
        H = -1j * collection_evaluate(ham_c, Hops, Hd)
    if Dstatic is None and Dops is None:
        return np.matmul(H, rho) - np.matmul(rho, H)
    A = 0.0
    if Dstatic is not None:
        Ds = np.asarray(Dstatic)
        A = A + (-0.5) * np.sum(np.matmul(np.swapaxes(Ds.conj(), -1, -2), Ds), axis=0)
    if Dops is not None:
        Do = np.asarray(Dops)
        A = A + np.tensordot(dis_c, -0.5 * np.matmul(np.swapaxes(Do.conj(), -1, -2), Do), axes=1)
    if H is not None:
        left = np.matmul(H + A, rho)
        right = np.matmul(rho, A - H)
    else:
        left = np.matmul(A, rho)
        right = np.matmul(rho, A)
    r = rho
    if rho.ndim == 3:
        r = rho[:, None, :, :]
    both = 0.0
    if Dstatic is not None:
        Ds = np.asarray(Dstatic)
        both = both + np.sum(np.matmul(Ds, np.matmul(r, np.swapaxes(Ds.conj(), -1, -2))), axis=-3)
    if Dops is not None:
        Do = np.asarray(Dops)
        mats = np.matmul(Do, np.matmul(r, np.swapaxes(Do.conj(), -1, -2)))
        both = both + np.tensordot(dis_c, mats.real, axes=(-1, -3)) \
            + 1j * np.tensordot(dis_c, mats.imag, axes=(-1, -3))
    return left + right + both


def vec_frame_phase(d):
    """mu with exp(-i mu t) the per-row pre-phase of a vectorised state: row a = i + k n of
    vec_F(rho) carries conj-conjugation e_i conj(e_k) -> mu_a = lam_i - lam_k."""
    lam = -np.imag(d)
    return (lam[:, None] - lam[None, :]).flatten(order="F")


# ----------------------------------------------------------------------------------------------
# a8  step grid
# ----------------------------------------------------------------------------------------------


def merge_t_args(t_span, t_eval=None):
    """solvers/solver_utils.py:46-93."""
    if t_eval is None:
        return np.asarray(t_span)
    t_span = np.array(t_span)
    t_eval = np.array(t_eval)
    if t_eval.ndim > 1:
        raise ValueError("t_eval must be 1 dimensional.")
    if np.min(t_eval) < np.min(t_span) or np.max(t_eval) > np.max(t_span):
        raise ValueError("t_eval entries must lie in t_span.")
    direction = np.sign(t_span[1] - t_span[0])
    if np.any(direction * np.diff(t_eval) < 0.0):
        raise ValueError("t_eval must be ordered according to the direction of integration.")
    return np.append(np.append(t_span[0], t_eval), t_span[1])


def fixed_step_sizes(t_span, t_eval, max_dt):
    """(t_list, h_list, n_steps_list); solvers/fixed_step_solvers.py:616-653."""
    t_list = np.array(merge_t_args(t_span, t_eval))
    max_dt = np.array(max_dt)
    delta = np.diff(t_list)
    n_steps = np.abs(delta / max_dt).astype(int)
    for i, (dt_i, n_i) in enumerate(zip(delta, n_steps)):
        if n_i == 0:
            n_steps[i] = 1
        elif np.abs(dt_i / n_i) / max_dt > 1 + 1e-15:
            n_steps[i] = n_i + 1
    return t_list, np.array(delta / n_steps), n_steps


# ----------------------------------------------------------------------------------------------
# a7 / a9  steppers + the template loop
# ----------------------------------------------------------------------------------------------


def rk4_step(rhs: Callable, t, y, h):
    """solvers/fixed_step_solvers.py:62-73 (note the div6 * h * (...) evaluation order)."""
    h2 = 0.5 * h
    tm = t + h2
    k1 = rhs(t, y)
    k2 = rhs(tm, y + h2 * k1)
    k3 = rhs(tm, y + h2 * k2)
    k4 = rhs(t + h, y + h * k3)
    return y + (1.0 / 6) * h * (k1 + 2 * k2 + 2 * k3 + k4)


def expm_step(generator: Callable, t, y, h):
    """Magnus order 1: expm(G(t + h/2) h) @ y; solvers/fixed_step_solvers.py:343-346,400-401."""
    return _scipy_expm(generator(t + (h / 2)) * h) @ y


def matrix_commutator(m1, m2):
    """solvers/fixed_step_solvers.py:314-324."""
    return m1 @ m2 - m2 @ m1


def magnus_propagator(generator: Callable, t0, h, magnus_order: int = 1, expm_func=None):
    """The one-step propagator of get_exponential_take_step for Magnus orders 1, 2 and 3
    (Gauss-Legendre nodes and the commutator-free-of-derivatives forms);
    solvers/fixed_step_solvers.py:327-395."""
    expm_func = _scipy_expm if expm_func is None else expm_func
    if magnus_order == 1:
        return expm_func(generator(t0 + (h / 2)) * h)
    if magnus_order == 2:
        c1 = 0.5 - np.sqrt(3) / 6
        c2 = 0.5 + np.sqrt(3) / 6
        p2 = np.sqrt(3) / 12
        g1 = generator(t0 + c1 * h)
        g2 = generator(t0 + c2 * h)
        return expm_func(h * (g1 + g2) / 2 + p2 * (h**2) * matrix_commutator(g2, g1))
    if magnus_order == 3:
        d1 = 0.5 - np.sqrt(15) / 10
        d2 = 0.5
        d3 = 0.5 + np.sqrt(15) / 10
        c0 = np.sqrt(15) / 3
        c1 = 10.0 / 3
        g1 = generator(t0 + d1 * h)
        g2 = generator(t0 + d2 * h)
        g3 = generator(t0 + d3 * h)
        a1 = h * g2
        a2 = c0 * h * (g3 - g1)
        a3 = c1 * h * (g3 - 2 * g2 + g1)
        comm1 = matrix_commutator(a1, a2)
        comm2 = matrix_commutator(2 * a3 + comm1, a1) / 60
        return expm_func(a1 + (a3 / 12) + matrix_commutator(-20 * a1 - a3 + comm1, a2 + comm2) / 240)
    raise ValueError("Only magnus_order 1, 2, and 3 are supported.")


def magnus_step(magnus_order: int):
    """take_step(generator, t0, y, h) = propagator @ y; solvers/fixed_step_solvers.py:397-401."""
    def take_step(generator, t0, y, h):
        return magnus_propagator(generator, t0, h, magnus_order) @ y
    return take_step


def magnus_node_offsets(magnus_order: int) -> np.ndarray:
    """Generator evaluation points of one step as fractions of h (solvers/fixed_step_solvers.py:346,350-351,
    367-369); the host code under test builds its time tables from the same expressions."""
    if magnus_order == 1:
        return np.array([0.5])
    if magnus_order == 2:
        return np.array([0.5 - np.sqrt(3) / 6, 0.5 + np.sqrt(3) / 6])
    if magnus_order == 3:
        return np.array([0.5 - np.sqrt(15) / 10, 0.5, 0.5 + np.sqrt(15) / 10])
    raise ValueError("Only magnus_order 1, 2, and 3 are supported.")


def fixed_step_solve(take_step, fn, t_span, y0, max_dt, t_eval=None):
    """solvers/fixed_step_solvers.py:441-459; returns (t, ys) with endpoints trimmed when
    t_eval is given (solvers/solver_utils.py:112-119)."""
    y0 = np.asarray(y0)
    t_list, h_list, n_list = fixed_step_sizes(t_span, t_eval, max_dt)
    ys = [y0]
    for t0, h, n in zip(t_list, h_list, n_list):
        y = ys[-1]
        t = t0
        for _ in range(n):
            y = take_step(fn, t, y, h)
            t = t + h
        ys.append(y)
    ys = np.asarray(ys)
    if t_eval is not None:
        return t_list[1:-1], ys[1:-1]
    return t_list, ys


def rk4_step_propagator(generator: Callable, t, h):
    """take_step of jax_RK4_parallel_solver (solvers/fixed_step_solvers.py:225-237): the RK4 step as a matrix."""
    ident = np.eye(np.asarray(generator(t)).shape[-1], dtype=complex)
    h2 = 0.5 * h
    gh2 = generator(t + h2)
    k1 = generator(t)
    k2 = gh2 @ (ident + h2 * k1)
    k3 = gh2 @ (ident + h2 * k2)
    k4 = generator(t + h) @ (ident + h * k3)
    return ident + (1.0 / 6) * h * (k1 + 2 * k2 + 2 * k3 + k4)


def parallel_solve(step_propagator: Callable, generator: Callable, t_span, y0, max_dt, t_eval=None):
    """fixed_step_lmde_solver_parallel_template_jax (solvers/fixed_step_solvers.py:524-613) without JAX: step times
    t + h * arange(n) per interval (not accumulated), one propagator per step, cumulative products (later steps on the
    left), results at the interval ends = cumulative propagator applied to y0.  step_propagator(generator, t, h)."""
    y0 = np.asarray(y0, dtype=complex)
    t_list, h_list, n_list = fixed_step_sizes(t_span, t_eval, max_dt)
    ys = [y0]
    total = None
    for t, h, n in zip(t_list, h_list, n_list):
        for ts in t + h * np.arange(n):
            P = step_propagator(generator, ts, h)
            total = P if total is None else P @ total
        ys.append(total @ y0)
    ys = np.asarray(ys)
    if t_eval is not None:
        return t_list[1:-1], ys[1:-1]
    return t_list, ys


def solve_hamiltonian_parallel(static_operator, operators, specs, frame_operator, t_span, y0, max_dt, kind="RK4",
                               magnus_order=1, t_eval=None, hamiltonian=True):
    """solve_lmde(model, method="jax_RK4_parallel" | "jax_expm_parallel") restated on NumPy (the reference runs these
    on JAX only, which is absent here: this restatement is pinned only through its agreement with the sequential
    solvers, which are pinned by the fixtures -- same propagators, associativity of the matrix product)."""
    Gd, G, d, U = generator_model_operators(static_operator, operators, frame_operator, hamiltonian)
    y0 = np.asarray(y0, dtype=complex)
    yfb = y0 if U is None else U.conj().T @ y0
    gen = lambda t: model_generator(t, specs, G, Gd, d)  # noqa: E731
    if kind == "RK4":
        step = rk4_step_propagator
    else:
        step = lambda g, t, h: magnus_propagator(g, t, h, magnus_order)  # noqa: E731
    t, ys = parallel_solve(step, gen, t_span, yfb, max_dt, t_eval)
    if U is not None:
        ys = (U @ ys.T).T if y0.ndim == 1 else U @ ys
    return t, ys


def stage_time_grid(t0, h, n_steps):
    """The 2S+1 stage times the template visits for one interval, with the same accumulation as
    the loop above (t <- t + h; midpoints t + 0.5 h): [t_0, t_0+h/2, t_1, t_1+h/2, ..., t_S]."""
    starts = np.cumsum(np.concatenate([[t0], np.full(n_steps, h)]))  # sequential, like the loop
    out = np.empty(2 * n_steps + 1)
    out[0::2] = starts
    out[1::2] = starts[:-1] + 0.5 * h
    return out


# ----------------------------------------------------------------------------------------------
# a11  solve_lmde end to end
# ----------------------------------------------------------------------------------------------


def solve_hamiltonian(static_operator, operators, specs, frame_operator, t_span, y0, max_dt,
                      method="RK4", t_eval=None, hamiltonian=True, magnus_order=1):
    """solve_lmde(HamiltonianModel/GeneratorModel, method in {"RK4","scipy_expm"}) for a model
    given out of the frame basis; solvers/solver_functions.py:315-327,342-371,396-448."""
    Gd, G, d, U = generator_model_operators(static_operator, operators, frame_operator, hamiltonian)
    y0 = np.asarray(y0, dtype=complex)
    yfb = y0 if U is None else U.conj().T @ y0
    if method == "RK4":
        t, ys = fixed_step_solve(rk4_step, lambda t, y: model_rhs(t, y, specs, G, Gd, d),
                                 t_span, yfb, max_dt, t_eval)
    elif method == "scipy_expm":
        step = expm_step if magnus_order == 1 else magnus_step(magnus_order)
        t, ys = fixed_step_solve(step, lambda t: model_generator(t, specs, G, Gd, d),
                                 t_span, yfb, max_dt, t_eval)
    else:
        raise ValueError(method)
    if U is not None:
        if y0.ndim == 1:
            ys = (U @ ys.T).T
        else:
            ys = U @ ys
    return t, ys


def solve_vectorized_lindblad(static_hamiltonian, hamiltonian_operators, ham_specs,
                              static_dissipators, dissipator_operators, dis_specs,
                              frame_operator, t_span, y0, max_dt, method="scipy_expm",
                              t_eval=None, magnus_order=1):
    """solve_lmde(LindbladModel(vectorized=True)); y0 is vec_F(rho) of shape (n^2,) or (n^2,B);
    lindblad_model.py:436-538, solver_functions.py:397-399,436-441."""
    Hd, Hops, Ds, Do, d, U = lindblad_model_operators(static_hamiltonian, hamiltonian_operators,
                                                      static_dissipators, dissipator_operators,
                                                      frame_operator)
    S, ops = vectorized_lindblad_collection(Hd, Hops, Ds, Do)

    def coeffs(t):
        parts = []
        if Hops is not None:
            parts.append(signal_list_values(ham_specs, t))
        if Do is not None:
            parts.append(signal_list_values(dis_specs, t))
        return None if not parts else np.concatenate(parts, axis=-1)

    def generator(t):
        G = collection_evaluate(coeffs(t), ops, S)
        return G if d is None else vectorized_map_into_frame(d, t, G)

    def rhs(t, y):
        return generator(t) @ y

    y0 = np.asarray(y0, dtype=complex)
    VU = None if U is None else np.kron(U.conj(), U)
    yfb = y0 if VU is None else VU.conj().T @ y0
    step = rk4_step if method == "RK4" else (expm_step if magnus_order == 1 else magnus_step(magnus_order))
    fn = rhs if method == "RK4" else generator
    t, ys = fixed_step_solve(step, fn, t_span, yfb, max_dt, t_eval)
    if VU is not None:
        ys = (VU @ ys.T).T if y0.ndim == 1 else VU @ ys
    return t, ys


# ----------------------------------------------------------------------------------------------
# sweep mode: per-column signals == the reference's sequential list-of-simulations loop
# (solvers/solver_classes.py:556-590): column b is solved with its own SignalList.
# ----------------------------------------------------------------------------------------------


def solve_hamiltonian_sweep(static_operator, operators, specs_per_column: List, frame_operator,
                            t_span, y0, max_dt, t_eval=None):
    y0 = np.asarray(y0, dtype=complex)
    outs = []
    t = None
    for b, specs in enumerate(specs_per_column):
        t, ys = solve_hamiltonian(static_operator, operators, specs, frame_operator, t_span,
                                  y0[:, b], max_dt, "RK4", t_eval)
        outs.append(ys)
    return t, np.stack(outs, axis=-1)


# ----------------------------------------------------------------------------------------------
# synthetic workloads of SURVEY.md section 8(d) (shared by golden generation, tests and bench)
# ----------------------------------------------------------------------------------------------


def herm(rng, n):
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    return (A + A.conj().T) / (2 * np.sqrt(n))


def synthetic_schrodinger(n, K, B, seed):
    """H0 = 5 herm, H_j = herm, unit-norm complex-normal columns, signals a_j=0.1(j+1),
    nu_j = 0.2 j + 0.05, phi_j = 0.3 j (SURVEY 8(d) / BASELINE.md section 3)."""
    rng = np.random.default_rng(seed)
    H0 = 5 * herm(rng, n)
    Hs = np.array([herm(rng, n) for _ in range(K)])
    Y = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    Y = Y / np.linalg.norm(Y, axis=0, keepdims=True)
    sig = [(0.1 * (j + 1), 0.2 * j + 0.05, 0.3 * j) for j in range(K)]
    return H0, Hs, Y, sig


def synthetic_lindblad(n, K, n_diss, B, seed):
    """cfg3: H0 = 5 herm(n), K herm drive ops, n_diss static dissipators 0.05 N(0,1) real,
    B random pure-state density matrices column-stacked."""
    rng = np.random.default_rng(seed)
    H0 = 5 * herm(rng, n)
    Hs = np.array([herm(rng, n) for _ in range(K)])
    Ls = 0.05 * rng.standard_normal((n_diss, n, n))
    psi = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    psi = psi / np.linalg.norm(psi, axis=0, keepdims=True)
    rho = np.einsum("ib,kb->ikb", psi, psi.conj())  # rho_b[i,k]
    Y = rho.reshape(n * n, B, order="F")  # vec_F: row index i + k n
    sig = [(0.1 * (j + 1), 0.2 * j + 0.05, 0.3 * j) for j in range(K)]
    return H0, Hs, Ls, Y, sig


# ----------------------------------------------------------------------------------------------
# f4  final-state measurement: out of frame -> dressed basis -> normalise ->
#     memory-slot outcome probabilities
# ----------------------------------------------------------------------------------------------


def dressed_state_decomposition(operator):
    """eigh sorted by overlap with the elementary basis; backend/backend_utils.py:31-80."""
    evals, evecs = np.linalg.eigh(np.array(operator))
    dressed_evals = np.zeros_like(evals)
    dressed_states = np.zeros_like(evecs)
    found = []
    for eigval, evec in zip(evals, evecs.transpose()):
        position = int(np.argmax(np.abs(evec)))
        if position in found:
            raise ValueError("Dressed-state sorting failed due to non-unique np.argmax(np.abs(evec)).")
        found.append(position)
        dressed_states[:, position] = evec
        dressed_evals[position] = eigval
    return dressed_evals, dressed_states


def subsystem_probabilities_dict(probs, dims, qargs):
    """Restatement of qiskit.quantum_info Statevector.probabilities_dict(qargs) -- the absent third-party
    dependency at this boundary (qiskit, unpinned in the reference's setup.py; algorithm of
    QuantumState._subsystem_probabilities + _vector2dict): reshape |amplitude|^2 to reversed(dims), sum the
    unmeasured subsystems, order the remaining axes so that qargs[0] is the LAST character of the key;
    zero-probability outcomes are dropped.  Called from backend/dynamics_backend.py:862."""
    dims = list(dims)
    ndim = len(dims)
    tens = np.reshape(np.asarray(probs, dtype=float), list(reversed(dims)))
    qaxes = [ndim - 1 - q for q in reversed(qargs)]
    sum_axis = tuple(i for i in range(ndim) if i not in qaxes)
    if sum_axis:
        tens = np.sum(tens, axis=sum_axis)
    perm = np.argsort(np.argsort(qaxes))
    tens = np.transpose(tens, axes=perm)
    flat = np.reshape(tens, (tens.size,))
    sub_dims = [dims[q] for q in qargs]
    out = {}
    for idx, p in enumerate(flat):
        if p == 0:
            continue
        digits, rem = [], idx
        for d in sub_dims:  # qargs[0] is the least significant digit
            digits.append(str(rem % d))
            rem //= d
        out["".join(reversed(digits))] = float(p)
    return out


def memory_slot_probabilities(probability_dict, memory_slot_indices, num_memory_slots=None, max_outcome_value=None):
    """backend/backend_utils.py:106-147."""
    num_memory_slots = num_memory_slots or (max(memory_slot_indices) + 1)
    out = {}
    for level_str, prob in probability_dict.items():
        result = ["0"] * num_memory_slots
        for idx, level in zip(memory_slot_indices, reversed(level_str)):
            if max_outcome_value and int(level) > max_outcome_value:
                level = str(max_outcome_value)
            result[-(idx + 1)] = level
        key = "".join(result)
        out[key] = out.get(key, 0.0) + prob
    return out


def final_state_memory_probabilities(y_frame, t, frame_operator, dressed_states, subsystem_dims, measurement_subsystems,
                                     memory_slot_indices, num_memory_slots=None, max_outcome_value=None, normalize=True):
    """One final state (standard basis, in the rotating frame) -> {memory-slot outcome: probability};
    backend/dynamics_backend.py:846-866 for a Statevector."""
    d, U = frame_decompose(frame_operator)
    y = np.asarray(y_frame, dtype=complex)
    if d is not None:
        yfb = y if U is None else U.conj().T @ y
        yfb = state_out_of_frame(d, t, yfb)
        y = yfb if U is None else U @ yfb
    y = dressed_states.conj().T @ y
    if normalize:
        y = y / np.linalg.norm(y)
    probs = np.abs(y) ** 2
    pd = subsystem_probabilities_dict(probs, subsystem_dims, measurement_subsystems)
    return memory_slot_probabilities(pd, memory_slot_indices, num_memory_slots, max_outcome_value)
