"""Worker for test_live_reference_cpu.py: the ORACLE against the live, unmodified reference on exactly the randomised
cases that tests/test_fuzz_gpu.py runs on the GPU (same seeds, same case generator) -- so that for every one of those cases
GPU == oracle (on the B200) and oracle == reference (here) are both checked."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle.ref_shim as shim  # noqa: E402

assert shim.reference_available()
import qiskit_dynamics as ref  # noqa: E402
from qiskit_dynamics.models import HamiltonianModel, LindbladModel  # noqa: E402
from qiskit_dynamics import solve_lmde  # noqa: E402

import test_fuzz_gpu as fz  # noqa: E402  (case generators only; nothing here touches a GPU)
from oracle import numpy_oracle as orc  # noqa: E402

TOL = 1e-11
worst = 0.0


def col_l2(a, b):
    d = np.asarray(a) - np.asarray(b)
    return float(np.max(np.linalg.norm(d.reshape(d.shape[0], -1) if d.ndim > 1 else d.reshape(-1, 1), axis=0)))


# ---- shared-signal solves (test_random_shared_signal_solves) ----
for case in range(24):
    rng = np.random.default_rng(1000 + case)
    n, K, B, frame, seed = fz.random_case(rng)
    discrete = bool(case % 3 == 0)
    H0, Hs, Y, fr, specs, sigs = fz.build(ref, n, K, B, frame, seed, rng, discrete)
    method = ["RK4", "scipy_expm"][case % 2]
    order = int(rng.integers(1, 4)) if method == "scipy_expm" else 1
    T = float(rng.uniform(0.05, 0.12))
    t_span = [0.0, T] if case % 5 else [T, 0.0]
    max_dt = T / float(rng.integers(3, 9)) * (1.0 if method == "scipy_expm" else 0.5)
    t_eval = None
    if case % 4 == 1:
        pts = np.sort(rng.uniform(min(t_span), max(t_span), 3))
        t_eval = pts if t_span[0] < t_span[1] else pts[::-1]
    y0 = Y[:, 0] if (B == 1 and case % 2) else Y
    model = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs, rotating_frame=fr)
    kw = dict(magnus_order=order) if method == "scipy_expm" else {}
    r = solve_lmde(model, t_span=t_span, y0=y0, method=method, max_dt=max_dt, t_eval=t_eval, **kw)
    t_o, y_o = orc.solve_hamiltonian(H0, Hs, specs, fr, t_span, y0, max_dt, method=method, t_eval=t_eval, **kw)
    assert np.array_equal(np.asarray(r.t), np.asarray(t_o)), case
    err = max(col_l2(np.asarray(r.y)[i], y_o[i]) for i in range(y_o.shape[0]))
    assert err < TOL, ("shared", case, err)
    worst = max(worst, err)

# ---- sweeps (test_random_sweeps_through_solver): per-simulation reference solves ----
for case in range(16):
    rng = np.random.default_rng(5000 + case)
    n, K, nsim, frame, seed = fz.random_case(rng)
    nsim = max(nsim, 2)
    discrete = bool(case % 2)
    H0, Hs, Y, fr, specs, sigs = fz.build(ref, n, K, 1, frame, seed, rng, discrete)
    scales = 0.4 + rng.uniform(0.0, 1.2, nsim)
    T = float(rng.uniform(0.04, 0.1))
    max_dt = T / float(rng.integers(4, 12))
    model = HamiltonianModel(static_operator=H0, operators=Hs, rotating_frame=fr)
    for b in (0, nsim - 1):  # first and last simulation of the list
        s = scales[b]
        if discrete:
            model.signals = [ref.DiscreteSignal(dt=x.dt, samples=s * np.asarray(x.samples), start_time=x.start_time,
                                                carrier_freq=x.carrier_freq, phase=x.phase) for x in sigs]
            sp = [orc.SigSpec(("discrete", q.envelope[1], s * q.envelope[2], q.envelope[3]), q.carrier_freq, q.phase) for q in specs]
        else:
            model.signals = [ref.Signal(s * q.envelope, q.carrier_freq, q.phase) for q in specs]
            sp = [orc.SigSpec(s * q.envelope, q.carrier_freq, q.phase) for q in specs]
        r = solve_lmde(model, t_span=[0.0, T], y0=Y[:, 0], method="RK4", max_dt=max_dt)
        _, yb = orc.solve_hamiltonian(H0, Hs, sp, fr, [0.0, T], Y[:, 0], max_dt)
        err = float(np.linalg.norm(np.asarray(r.y)[-1] - yb[-1]))
        assert err < TOL, ("sweep", case, b, err)
        worst = max(worst, err)

# ---- Lindblad (test_random_lindblad_solves) ----
for case in range(12):
    rng = np.random.default_rng(9000 + case)
    n = int(rng.choice([2, 3, 4, 5, 6]))
    K = int(rng.integers(1, 4))
    B = int(rng.choice([1, 3, 8, 11]))
    nd = int(rng.integers(2, 5))
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, nd, B, int(rng.integers(1, 10**6)))
    Ls = 4.0 * Ls
    split = int(rng.integers(0, nd + 1))
    Lstat = Ls[:split] if split > 0 else None
    Ldyn = (Ls[split:] + 0.3j * Ls[split:][::-1]) if split < nd else None
    dsig = [(0.5 + 0.2 * j, 0.07 * j, 0.3 * j) for j in range(nd - split)]
    frame = [None, H0, np.diag(H0).real][case % 3]
    method = ["scipy_expm", "RK4"][case % 2]
    order = int(rng.integers(1, 4)) if method == "scipy_expm" else 1
    T = float(rng.uniform(0.1, 0.3))
    max_dt = T / float(rng.integers(3, 8)) * (1.0 if method == "scipy_expm" else 0.25)
    sp = [orc.SigSpec(*s) for s in sig]
    dsp = [orc.SigSpec(*s) for s in dsig] if Ldyn is not None else None
    kw = dict(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=[ref.Signal(*s) for s in sig],
              static_dissipators=Lstat, dissipator_operators=Ldyn,
              dissipator_signals=[ref.Signal(*s) for s in dsig] if Ldyn is not None else None, rotating_frame=frame)
    extra = dict(magnus_order=order) if method == "scipy_expm" else {}
    r = solve_lmde(LindbladModel(vectorized=True, **kw), t_span=[0, T], y0=Y, method=method, max_dt=max_dt, **extra)
    _, ys = orc.solve_vectorized_lindblad(H0, Hs, sp, Lstat, Ldyn, dsp, frame, [0, T], Y, max_dt, method, **extra)
    err = col_l2(np.asarray(r.y)[-1], ys[-1])
    assert err < TOL, ("lindblad", case, err)
    worst = max(worst, err)
    if method == "RK4":
        rho = np.array([Y[:, b].reshape(n, n, order="F") for b in range(B)])
        rm = solve_lmde(LindbladModel(vectorized=False, **kw), t_span=[0, T], y0=rho, method="RK4", max_dt=max_dt)
        vec = np.stack([m_.flatten(order="F") for m_ in np.asarray(rm.y)[-1]], axis=-1)
        assert col_l2(vec, ys[-1]) < TOL, ("lindblad matrix form", case)
print(f"LIVE_REFERENCE_FUZZ_OK cases=52 worst_col_l2={worst:.2e}")
