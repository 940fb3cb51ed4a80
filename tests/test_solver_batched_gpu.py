"""Row f1 widened: Solver.solve on LISTS of simulations -- the reference's sequential loop
(solvers/solver_classes.py:556-590, argument expansion solvers/solver_utils.py:230-287) -- as batched launches.
The reference's own contract (test/dynamics/solvers/test_solver_classes.py:1388-1599) is `results[i] == the i-th
individual solve`; here each batched route is compared with individual solves (our kernels, one simulation at a
time) and with the oracle.  Tolerance 1e-10."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import bench_workloads as W  # noqa: E402
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
    return x.detach().cpu().numpy()


def test_state_ensemble_shares_one_signal_set(qd):
    """A list of y0 with ONE signal specification runs on the shared-signal kernels (one solve, states as columns) for
    RK4, the exponential stepper with magnus_order, and the time-parallel solver; matrices as initial states too."""
    n, K, B = 24, 3, 7
    H0, Hs, Y, sig = W.schrodinger(n, K, B, 5)
    solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, rotating_frame=H0)
    sigs = [qd.Signal(a, nu, ph) for a, nu, ph in sig]
    specs = [orc.SigSpec(a, nu, ph) for a, nu, ph in sig]
    y0_list = [Y[:, b].copy() for b in range(B)]
    before = qd._abi.launch_count()
    out = solver.solve(t_span=[0, 0.3], y0=y0_list, signals=sigs, method="RK4", max_dt=1e-2)
    batched_launches = qd._abi.launch_count() - before
    assert len(out) == B and out[0].y.shape == (2, n)
    _, ys = orc.solve_hamiltonian(H0, Hs, specs, H0, [0, 0.3], Y, 1e-2)
    assert max_col_l2(np.stack([npy(r.y[-1]) for r in out], axis=-1), ys[-1]) < TOL
    before = qd._abi.launch_count()
    one = solver.solve(t_span=[0, 0.3], y0=y0_list[3], signals=sigs, method="RK4", max_dt=1e-2)
    assert batched_launches <= (qd._abi.launch_count() - before) + 6  # the whole list costs what one simulation costs (+ one-time operand packing)
    assert float((one.y[-1] - out[3].y[-1]).abs().max()) < 1e-13
    # exponential stepper at Magnus order 2 with t_eval, and the time-parallel solver: keywords pass through
    oe = solver.solve(t_span=[0, 0.3], y0=y0_list, signals=sigs, method="scipy_expm", max_dt=0.05, magnus_order=2, t_eval=[0.1, 0.3])
    _, ye = orc.solve_hamiltonian(H0, Hs, specs, H0, [0, 0.3], Y, 0.05, method="scipy_expm", magnus_order=2, t_eval=[0.1, 0.3])
    assert oe[0].y.shape == (2, n)
    for i in range(2):
        assert max_col_l2(np.stack([npy(r.y[i]) for r in oe], axis=-1), ye[i]) < TOL
    op = solver.solve(t_span=[0, 0.3], y0=y0_list, signals=sigs, method="jax_RK4_parallel", max_dt=1e-2)
    assert max_col_l2(np.stack([npy(r.y[-1]) for r in op], axis=-1), ys[-1]) < 1e-9
    # matrices as initial states (n, m): concatenated columns
    mats = [Y[:, :3].copy(), Y[:, 2:5].copy(), Y[:, 4:7].copy()]
    om = solver.solve(t_span=[0, 0.3], y0=mats, signals=sigs, method="RK4", max_dt=1e-2)
    assert om[1].y.shape == (2, n, 3)
    assert max_col_l2(npy(om[1].y[-1]), ys[-1][:, 2:5]) < TOL and max_col_l2(npy(om[2].y[-1]), ys[-1][:, 4:7]) < TOL


def test_sweep_with_matrix_initial_states(qd):
    """Per-simulation signals AND an (n, m) matrix of initial states per simulation: every simulation is a group of m
    columns with its own signal values, one sweep launch."""
    n, K, nsim, m = 16, 4, 6, 3
    H0, Hs, Y, sig = W.schrodinger(n, K, nsim * m, 9)
    solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, rotating_frame=np.diag(H0).real)
    lists = [[qd.Signal(a * (0.5 + b / nsim), nu + 0.01 * b, ph) for a, nu, ph in sig] for b in range(nsim)]
    y0s = [Y[:, b * m:(b + 1) * m].copy() for b in range(nsim)]
    before = qd._abi.launch_count()
    out = solver.solve(t_span=[0, 0.2], y0=y0s, signals=lists, method="RK4", max_dt=5e-3)
    assert qd._abi.launch_count() - before < 15
    for b in range(nsim):
        specs = [orc.SigSpec(a * (0.5 + b / nsim), nu + 0.01 * b, ph) for a, nu, ph in sig]
        _, ys = orc.solve_hamiltonian(H0, Hs, specs, np.diag(H0).real, [0, 0.2], y0s[b], 5e-3)
        assert out[b].y.shape == (2, n, m)
        assert max_col_l2(npy(out[b].y[-1]), ys[-1]) < TOL


def test_vectorised_lindblad_sweep_beyond_256(qd):
    """dim 17 -> vec-rho 289 > 256: per-simulation signals used to fall back to the sequential loop; now the generic
    sweep path (one DMMA GEMM per operator and stage, signal values as column scales) runs the whole list at once."""
    n, K, nsim = 17, 2, 5
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 3, nsim, 77)
    solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, static_dissipators=Ls, rotating_frame=np.diag(H0).real,
                       vectorized=True)
    lists = [[qd.Signal(a * (1 + 0.3 * b), nu, ph + 0.1 * b) for a, nu, ph in sig] for b in range(nsim)]
    y0s = [Y[:, b].copy() for b in range(nsim)]
    before = qd._abi.launch_count()
    out = solver.solve(t_span=[0, 0.05], y0=y0s, signals=lists, method="RK4", max_dt=5e-3)
    batched = qd._abi.launch_count() - before
    assert len(out) == nsim and out[0].y.shape == (2, n * n)
    for b in range(nsim):
        specs = [orc.SigSpec(a * (1 + 0.3 * b), nu, ph + 0.1 * b) for a, nu, ph in sig]
        _, ys = orc.solve_vectorized_lindblad(H0, Hs, specs, Ls, None, None, np.diag(H0).real, [0, 0.05], Y[:, b], 5e-3, method="RK4")
        assert np.linalg.norm(npy(out[b].y[-1]) - ys[-1]) < TOL
    # the sequential route (one simulation at a time) gives the same states with nsim times the launches
    before = qd._abi.launch_count()
    seq = [solver.solve(t_span=[0, 0.05], y0=y0s[b], signals=lists[b], method="RK4", max_dt=5e-3) for b in range(nsim)]
    assert qd._abi.launch_count() - before > batched
    assert max(float((seq[b].y[-1] - out[b].y[-1]).abs().max()) for b in range(nsim)) < 1e-12


def test_density_matrix_ensemble_non_vectorised(qd):
    """A list of (n, n) density matrices with one signal set on a non-vectorised LindbladModel: one (l, n, n) batch through
    the fused Lindblad kernel."""
    n, K, nsim = 9, 2, 6
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 3, nsim, 13)
    sigs = [qd.Signal(a, nu, ph) for a, nu, ph in sig]
    solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, static_dissipators=Ls, rotating_frame=H0)
    rhos = [Y[:, b].reshape(n, n, order="F").copy() for b in range(nsim)]
    before = qd._abi.launch_count()
    out = solver.solve(t_span=[0, 0.1], y0=rhos, signals=sigs, method="RK4", max_dt=2e-3)
    batched = qd._abi.launch_count() - before
    assert len(out) == nsim and out[0].y.shape == (2, n, n)
    specs = [orc.SigSpec(a, nu, ph) for a, nu, ph in sig]
    _, yv = orc.solve_vectorized_lindblad(H0, Hs, specs, Ls, None, None, H0, [0, 0.1], Y, 2e-3, method="RK4")
    for b in range(nsim):
        assert np.linalg.norm(npy(out[b].y[-1]).reshape(-1, order="F") - yv[-1][:, b]) < TOL
    before = qd._abi.launch_count()
    solver.solve(t_span=[0, 0.1], y0=rhos[0], signals=sigs, method="RK4", max_dt=2e-3)
    assert batched <= (qd._abi.launch_count() - before) + 8
