"""Randomised parity: the public API against the oracle over random shapes, frames, steppers, Magnus orders, t_eval
grids, integration directions and signal kinds (GPU).  Seeds are fixed, so a failure names a reproducible case; the
oracle is pinned to the reference by tests/test_oracle_golden.py.  Tolerance: 1e-10 on the north-star metric
(max column L2 error; bar 1e-8)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import max_col_l2  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402

TOL = 1e-10


@pytest.fixture(scope="module")
def qd():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import qiskit_dynamics_b200 as q
    q._abi.lib()
    return q


def npy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def random_case(rng):
    n = int(rng.choice([1, 2, 3, 5, 7, 8, 9, 12, 16, 17, 24, 31, 32, 33, 40, 48, 57, 64, 70]))
    K = int(rng.integers(1, 10))
    B = int(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 33]))
    frame = str(rng.choice(["none", "diag", "full"]))
    seed = int(rng.integers(1, 10**6))
    return n, K, B, frame, seed


def build(qd, n, K, B, frame, seed, rng, discrete):
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, seed)
    fr = {"none": None, "diag": np.diag(H0).real, "full": H0}[frame]
    if discrete:  # piecewise-constant envelopes (device signal-table route), some with a start offset
        dt = 0.013
        specs, sigs = [], []
        for j, (a, nu, ph) in enumerate(sig):
            samples = a * (rng.standard_normal(9) + 1j * rng.standard_normal(9))
            t0 = 0.0 if j % 2 == 0 else 0.02
            specs.append(orc.SigSpec(("discrete", dt, samples, t0), nu, ph))
            sigs.append(qd.DiscreteSignal(dt=dt, samples=samples, start_time=t0, carrier_freq=nu, phase=ph))
    else:
        specs = [orc.SigSpec(a, nu, ph) for (a, nu, ph) in sig]
        sigs = [qd.Signal(a, nu, ph) for (a, nu, ph) in sig]
    return H0, Hs, Y, fr, specs, sigs


@pytest.mark.parametrize("case", range(24))
def test_random_shared_signal_solves(qd, case):
    rng = np.random.default_rng(1000 + case)
    n, K, B, frame, seed = random_case(rng)
    discrete = bool(case % 3 == 0)
    H0, Hs, Y, fr, specs, sigs = build(qd, n, K, B, frame, seed, rng, discrete)
    method = ["RK4", "scipy_expm"][case % 2]
    order = int(rng.integers(1, 4)) if method == "scipy_expm" else 1
    T = float(rng.uniform(0.05, 0.12))
    t_span = [0.0, T] if case % 5 else [T, 0.0]  # every fifth case integrates backwards
    max_dt = T / float(rng.integers(3, 9)) * (1.0 if method == "scipy_expm" else 0.5)
    t_eval = None
    if case % 4 == 1:
        pts = np.sort(rng.uniform(min(t_span), max(t_span), 3))
        t_eval = pts if t_span[0] < t_span[1] else pts[::-1]
    y0 = Y[:, 0] if (B == 1 and case % 2) else Y
    model = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs, rotating_frame=fr)
    kw = dict(magnus_order=order) if method == "scipy_expm" else {}
    res = qd.solve_lmde(model, t_span=t_span, y0=y0, method=method, max_dt=max_dt, t_eval=t_eval, **kw)
    t_ref, y_ref = orc.solve_hamiltonian(H0, Hs, specs, fr, t_span, y0, max_dt, method=method, t_eval=t_eval,
                                         **({"magnus_order": order} if method == "scipy_expm" else {}))
    assert np.array_equal(np.asarray(res.t), np.asarray(t_ref))
    got = npy(res.y)
    assert got.shape == y_ref.shape, (n, K, B, frame, method, order)
    err = max(max_col_l2(got[i], y_ref[i]) for i in range(got.shape[0]))
    assert err < TOL, (n, K, B, frame, method, order, discrete, err)


@pytest.mark.parametrize("case", range(16))
def test_random_sweeps_through_solver(qd, case):
    """Lists of simulations through Solver.solve (one sweep-mode launch: formed-generator or operator-pass kernels,
    device signal tables when every term is a sampled / constant-envelope signal) == per-simulation oracle solves."""
    rng = np.random.default_rng(5000 + case)
    n, K, nsim, frame, seed = random_case(rng)
    nsim = max(nsim, 2)
    discrete = bool(case % 2)
    H0, Hs, Y, fr, specs, sigs = build(qd, n, K, 1, frame, seed, rng, discrete)
    scales = 0.4 + rng.uniform(0.0, 1.2, nsim)
    sig_lists, spec_lists = [], []
    for s in scales:
        if discrete:
            sig_lists.append([qd.DiscreteSignal(dt=x.dt, samples=s * npy(x.samples), start_time=x.start_time,
                                                carrier_freq=x.carrier_freq, phase=x.phase) for x in sigs])
            spec_lists.append([orc.SigSpec(("discrete", sp.envelope[1], s * sp.envelope[2], sp.envelope[3]), sp.carrier_freq, sp.phase)
                               for sp in specs])
        else:
            sig_lists.append([qd.Signal(s * sp.envelope, sp.carrier_freq, sp.phase) for sp in specs])
            spec_lists.append([orc.SigSpec(s * sp.envelope, sp.carrier_freq, sp.phase) for sp in specs])
    T = float(rng.uniform(0.04, 0.1))
    max_dt = T / float(rng.integers(4, 12))
    solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, rotating_frame=fr)
    out = solver.solve(t_span=[0.0, T], y0=Y[:, 0], signals=sig_lists, method="RK4", max_dt=max_dt)
    assert len(out) == nsim
    for b in range(nsim):
        _, yb = orc.solve_hamiltonian(H0, Hs, spec_lists[b], fr, [0.0, T], Y[:, 0], max_dt)
        err = float(np.linalg.norm(npy(out[b].y[-1]) - yb[-1]))
        assert err < TOL, (n, K, nsim, frame, discrete, b, err)


@pytest.mark.parametrize("case", range(12))
def test_random_lindblad_solves(qd, case):
    """Vectorised Lindblad models (row a10) with random term combinations, frames, steppers and Magnus orders against the
    oracle; and the non-vectorised model (row f2, batch of density matrices leading) against the vectorised one."""
    rng = np.random.default_rng(9000 + case)
    n = int(rng.choice([2, 3, 4, 5, 6]))
    K = int(rng.integers(1, 4))
    B = int(rng.choice([1, 3, 8, 11]))
    nd = int(rng.integers(2, 5))
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, nd, B, int(rng.integers(1, 10**6)))
    Ls = 4.0 * Ls  # make the dissipation visible over the short interval
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
    kw = dict(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=[qd.Signal(*s) for s in sig],
              static_dissipators=Lstat, dissipator_operators=Ldyn,
              dissipator_signals=[qd.Signal(*s) for s in dsig] if Ldyn is not None else None, rotating_frame=frame)
    mv = qd.LindbladModel(vectorized=True, **kw)
    extra = dict(magnus_order=order) if method == "scipy_expm" else {}
    res = qd.solve_lmde(mv, t_span=[0, T], y0=Y, method=method, max_dt=max_dt, **extra)
    _, ys = orc.solve_vectorized_lindblad(H0, Hs, sp, Lstat, Ldyn, dsp, frame, [0, T], Y, max_dt, method, **extra)
    err = max_col_l2(npy(res.y[-1]), ys[-1])
    assert err < TOL, (n, K, B, split, nd, method, order, err)
    if method == "RK4":
        mm = qd.LindbladModel(vectorized=False, **kw)
        rho = np.array([Y[:, b].reshape(n, n, order="F") for b in range(B)])
        rm = qd.solve_lmde(mm, t_span=[0, T], y0=rho, method="RK4", max_dt=max_dt)
        vec = np.stack([m_.flatten(order="F") for m_ in npy(rm.y[-1])], axis=-1)
        assert max_col_l2(vec, ys[-1]) < TOL
        # trace preservation of the Lindblad flow
        assert np.max(np.abs(np.trace(npy(rm.y[-1]), axis1=1, axis2=2) - 1.0)) < 1e-9
