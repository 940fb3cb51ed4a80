"""GPU parity of the drop-in surface (Signal / RotatingFrame / HamiltonianModel / LindbladModel /
solve_lmde / Solver) against fixtures produced by the UNMODIFIED reference (tests/golden/*.npz).

These read like the reference's own tests (test_generator_model.py, test_lindblad_model.py,
test_rotating_frame.py, test_solver_functions.py): build the model from the same arrays, call the
same methods, compare.  Tolerance: 1e-10 absolute on O(1) data (the north-star bar is 1e-8).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import MEASUREMENT_CASES, load_golden, max_col_l2  # noqa: E402
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


def close(a, b, tol=TOL):
    np.testing.assert_allclose(npy(a), np.asarray(b), rtol=0, atol=tol)


def sigs(qd, spec):
    return [qd.Signal(a, nu, ph) for (a, nu, ph) in np.asarray(spec)]


def test_operator_collection(qd):
    g = load_golden("collection")
    full = qd.OperatorCollection(static_operator=g["stat"], operators=g["ops"])
    nostat = qd.OperatorCollection(operators=g["ops"])
    onlystat = qd.OperatorCollection(static_operator=g["stat"])
    close(full.evaluate(g["c_real"]), g["eval_real"])
    close(full.evaluate(g["c_cplx"]), g["eval_cplx"])
    close(nostat.evaluate(g["c_real"]), g["eval_nostat"])
    close(onlystat.evaluate(None), g["eval_onlystat"])
    close(full.evaluate_rhs(g["c_real"], g["yv"]), g["rhs_v"])
    close(full.evaluate_rhs(g["c_real"], g["ym"]), g["rhs_m"])
    close(full.evaluate_rhs(g["c_cplx"], g["ym"]), g["rhs_m_cplx"])
    close(nostat(g["c_real"], g["ym"]), g["rhs_m_nostat"])
    close(onlystat(None, g["ym"]), g["rhs_m_onlystat"])
    assert full.dim == 6
    with pytest.raises(qd.QiskitError):
        qd.OperatorCollection().evaluate(None)
    with pytest.raises(qd.QiskitError):
        qd.OperatorCollection(operators=g["ops"], array_library="scipy_sparse")
    # explicit-loop check with 32 operators of 128x128 and complex coefficients
    # (reference test_operator_collections.py:82-94)
    rng = np.random.default_rng(342)
    ops = rng.uniform(-1, 1, (32, 128, 128)) + 1j * rng.uniform(-1, 1, (32, 128, 128))
    c = rng.uniform(-1, 1, 32) + 1j * rng.uniform(-1, 1, 32)
    close(qd.OperatorCollection(operators=ops).evaluate(c), sum(ci * o for ci, o in zip(c, ops)), 1e-11)


def test_rotating_frame(qd):
    g = load_golden("frame")
    rf = qd.RotatingFrame(g["H"])
    t = float(g["t"])
    y, op, sop = g["y"], g["op"], g["sop"]
    close(rf.frame_diag, g["frame_diag"])
    close(qd.RotatingFrame(-1j * g["H"]).frame_diag, g["frame_diag_anti"])
    rf1 = qd.RotatingFrame(g["diag1d"])
    close(rf1.frame_diag, g["frame_diag_1d"])
    assert rf1.frame_basis is None and rf.dim == 5
    close(rf.state_into_frame(t, y, y_in_frame_basis=True, return_in_frame_basis=True), g["into_fb"])
    close(rf.state_out_of_frame(t, y, y_in_frame_basis=True, return_in_frame_basis=True), g["outof_fb"])
    close(rf.state_into_frame(t, y), g["into_full"])
    close(rf.state_out_of_frame(t, y), g["outof_full"])
    close(rf.state_into_frame(t, y[:, 0]), g["into_full"][:, 0])
    close(rf.operator_into_frame(t, op, operator_in_frame_basis=True, return_in_frame_basis=True), g["op_into_fb"])
    close(rf.operator_into_frame(t, op), g["op_into_full"])
    close(rf.operator_out_of_frame(t, op), g["op_outof_full"])
    close(rf.generator_into_frame(t, op), g["gen_into_full"])
    close(rf.generator_out_of_frame(t, op), g["gen_outof_full"])
    close(rf.vectorized_map_into_frame(t, sop, operator_in_frame_basis=True, return_in_frame_basis=True), g["vec_into_fb"])
    close(rf.vectorized_map_into_frame(t, sop), g["vec_into_full"], 1e-9)
    close(rf.state_out_of_frame_basis(rf.state_into_frame_basis(y)), y)
    close(rf.operator_out_of_frame_basis(rf.operator_into_frame_basis(op)), op)
    close(rf1.state_into_frame(t, y), g["into_1d"])
    close(rf1.operator_into_frame(t, op), g["op_into_1d"])
    # stack of operators and the vectorised-operator convention (dim^2, k)
    stack = np.stack([op, 2 * op])
    close(rf.operator_into_frame(t, stack), np.stack([g["op_into_full"], 2 * g["op_into_full"]]))
    vec = np.stack([op.flatten(order="F"), 2 * op.flatten(order="F")], axis=1)
    out = npy(rf.operator_into_frame(t, vec, vectorized_operators=True))
    close(out[:, 0].reshape(5, 5, order="F"), g["op_into_full"])
    # null frame is the identity map
    rf0 = qd.RotatingFrame(None)
    close(rf0.state_into_frame(t, y), y)
    close(rf0.generator_into_frame(t, op), op)
    with pytest.raises(qd.QiskitError):
        qd.RotatingFrame(np.array([[1.0, 2.0], [3.0, 4.0]]))


def test_hamiltonian_model(qd):
    g = load_golden("hamiltonian_model")
    H0, Hs, Y, sig = g["H0"], g["Hs"], g["Y"], g["sig"]
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig), rotating_frame=frame)
        assert m.dim == 8
        for i, t in enumerate(g["ts"]):
            close(m(t, Y), g[f"rhs_{frame_name}_fb0_{i}"])
            close(m(t), g[f"gen_{frame_name}_fb0_{i}"])
        close(m(0.37, Y[:, 0]), g[f"rhsvec_{frame_name}_fb0"])
        if frame_name != "full":  # frame-basis quantities are eigenvector-phase dependent for a full frame
            m.in_frame_basis = True
            for i, t in enumerate(g["ts"]):
                close(m(t, Y), g[f"rhs_{frame_name}_fb1_{i}"])
                close(m(t), g[f"gen_{frame_name}_fb1_{i}"])
            close(m._operator_collection.operators, g[f"ops_{frame_name}"])
        # public operator accessors undo the -i fold (hamiltonian_model.py:134-150)
        m.in_frame_basis = False
        close(m.operators, Hs)
    # batch of columns == per-column (reference test_generator_model.py:615-674)
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig), rotating_frame=H0)
    full = npy(m(0.37, Y))
    for b in range(Y.shape[1]):
        close(m(0.37, Y[:, b]), full[:, b], 1e-13)
    # frame only, no static operator
    m = qd.HamiltonianModel(operators=Hs, signals=sigs(qd, sig), rotating_frame=H0)
    Gd_o, G_o, d_o, U_o = orc.generator_model_operators(None, Hs, H0)
    sp = [orc.SigSpec(a, nu, ph) for (a, nu, ph) in sig]
    close(m(0.37, Y), U_o @ orc.model_rhs(0.37, U_o.conj().T @ Y, sp, G_o, Gd_o, d_o))
    close(torch.diag(m._operator_collection.static_operator), np.diag(g["stat_nostatic"]))
    # GeneratorModel with a non-Hermitian generator in an anti-Hermitian frame
    gm = qd.GeneratorModel(static_operator=g["Gd"], operators=g["Gs"], signals=sigs(qd, sig), rotating_frame=-1j * H0)
    close(gm(0.37, Y), g["genmodel_rhs"], 1e-9)
    close(gm(0.37), g["genmodel_gen"], 1e-9)
    # error conventions
    with pytest.raises(qd.QiskitError):
        qd.HamiltonianModel(static_operator=np.array([[0, 1], [0, 0]]))
    with pytest.raises(qd.QiskitError):
        qd.HamiltonianModel()
    with pytest.raises(qd.QiskitError):
        qd.HamiltonianModel(operators=Hs, signals=sigs(qd, sig)[:2])
    with pytest.raises(qd.QiskitError):
        qd.HamiltonianModel(operators=Hs)(0.1, Y)  # no signals


def test_rk4_solves(qd):
    g = load_golden("rk4_solves")
    # cfg1 through solve_lmde and through Solver (plumbing)
    m = qd.HamiltonianModel(static_operator=g["cfg1_H0"], operators=[g["cfg1_H1"]], signals=[qd.Signal(1.0, 5.0)],
                            rotating_frame=g["cfg1_H0"])
    r = qd.solve_lmde(m, t_span=[0, 10.0], y0=g["cfg1_y0"], method="RK4", max_dt=1e-3)
    assert tuple(r.y.shape) == (2, 4) and list(r.t) == [0, 10.0]
    close(r.y, g["cfg1_y"], 1e-9)
    assert m.in_frame_basis is False
    s = qd.Solver(static_hamiltonian=g["cfg1_H0"], hamiltonian_operators=[g["cfg1_H1"]], rotating_frame=g["cfg1_H0"])
    r2 = s.solve(t_span=[0, 10.0], y0=g["cfg1_y0"], signals=[qd.Signal(1.0, 5.0)], method="RK4", max_dt=1e-3)
    close(r2.y, g["cfg1_solver_y"], 1e-9)
    assert s.model.signals is None
    # cfg4-like
    H0, Hs, Y, sig = orc.synthetic_schrodinger(128, 8, 8, 2004)
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig), rotating_frame=H0)
    r = qd.solve_lmde(m, t_span=[0, 0.05], y0=Y, method="RK4", max_dt=1e-3)
    assert max_col_l2(npy(r.y[-1]), g["cfg4_y"]) < TOL
    r = qd.solve_lmde(m, t_span=[0, 0.02], y0=Y, method="RK4", max_dt=1e-3, t_eval=[0.0, 0.005, 0.0125, 0.02])
    assert np.array_equal(r.t, g["cfg4_teval_t"])
    close(r.y, g["cfg4_teval_y"])
    close(qd.solve_lmde(m, t_span=[0, 1e-3], y0=Y, method="RK4", max_dt=1e-3).y[-1], g["cfg4_onestep_y"], 1e-12)
    close(qd.solve_lmde(m, t_span=[0.02, 0.0], y0=Y, method="RK4", max_dt=1e-3).y[-1], g["cfg4_back_y"])
    close(qd.solve_ode(m, t_span=[0, 0.05], y0=Y, method="RK4", max_dt=1e-3).y[-1], g["cfg4_y"])
    # cfg2-like sweep through Solver with a list of signal lists (one launch for the whole list)
    H0, Hs, Y, sig = orc.synthetic_schrodinger(32, 8, 1, 2002)
    B = 16
    s = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, rotating_frame=H0)
    sig_lists = [[qd.Signal(a * (0.5 + b / B), nu, ph) for (a, nu, ph) in sig] for b in range(B)]
    before = qd._abi.launch_count()
    res = s.solve(t_span=[0, 0.1], y0=Y[:, 0], signals=sig_lists, method="RK4", max_dt=1e-3)
    launches = qd._abi.launch_count() - before
    assert isinstance(res, list) and len(res) == B
    assert launches <= 8, f"sweep should be a handful of launches, saw {launches}"
    got = np.stack([npy(r.y[-1]) for r in res], axis=-1)
    assert max_col_l2(got, g["cfg2_y"]) < TOL
    # ... and the sequential path (different y0 shapes disable batching) agrees
    res1 = s.solve(t_span=[0, 0.1], y0=Y[:, 0], signals=sig_lists[3], method="RK4", max_dt=1e-3)
    close(res1.y[-1], g["cfg2_y"][:, 3])
    # odd dimension
    H0, Hs, Y, sg = g["odd_H0"], g["odd_Hs"], g["odd_Y"], g["odd_sig"]
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sg))
    close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y, method="RK4", max_dt=0.01).y[-1], g["odd_noframe_y"])
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sg), rotating_frame=np.diag(H0).real)
    close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y, method="RK4", max_dt=0.01).y[-1], g["odd_diagframe_y"])
    close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y[:, 0], method="RK4", max_dt=0.01).y[-1], g["odd_vec_y"])
    close(qd.solve_lmde(m, t_span=[0, 0.5], y0=np.eye(5, dtype=complex), method="RK4", max_dt=0.01).y[-1], g["odd_eye_y"])
    close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.01).y[-1], g["odd_expm_y"])
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sg), rotating_frame=H0)
    close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.01).y[-1], g["odd_expm_fullframe_y"])
    # DiscreteSignal drive, every stage time on a bin edge
    dt = float(g["disc_dt"])
    ds = [qd.DiscreteSignal(dt=dt, samples=g["disc_samples"][j], carrier_freq=0.3 * (j + 1), phase=0.1 * j) for j in range(2)]
    m = qd.HamiltonianModel(static_operator=g["disc_H0"], operators=g["disc_Hs"], signals=ds, rotating_frame=g["disc_H0"])
    close(qd.solve_lmde(m, t_span=[0, 2.0], y0=g["disc_Y"], method="RK4", max_dt=dt / 2).y[-1], g["disc_y"])
    m = qd.HamiltonianModel(static_operator=g["disc_H0"], operators=g["disc_Hs"], signals=ds)
    close(qd.solve_lmde(m, t_span=[0, 2.0], y0=g["disc_Y"], method="RK4", max_dt=dt / 2).y[-1], g["disc_noframe_y"])
    # callable generator / rhs (reference test_fixed_step_solvers.py style)
    rng = np.random.default_rng(3)
    A = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    Ad = qd.asarray(A)
    r = qd.solve_lmde(lambda t: Ad * np.cos(t), t_span=[0, 1.0], y0=np.eye(5, dtype=complex), method="RK4", max_dt=0.01)
    yref = orc.fixed_step_solve(orc.rk4_step, lambda t, y: (A * np.cos(t)) @ y, [0, 1.0], np.eye(5, dtype=complex), 0.01)[1]
    close(r.y, yref)
    r = qd.solve_lmde(lambda t: Ad * np.cos(t), t_span=[0, 1.0], y0=np.eye(5, dtype=complex), method="scipy_expm", max_dt=0.1)
    yref = orc.fixed_step_solve(orc.expm_step, lambda t: A * np.cos(t), [0, 1.0], np.eye(5, dtype=complex), 0.1)[1]
    close(r.y, yref, 1e-9)
    # error conventions
    with pytest.raises(qd.QiskitError):
        qd.solve_lmde(m, t_span=[0, 1], y0=g["disc_Y"], method="DOP853")
    with pytest.raises(qd.QiskitError):
        qd.solve_lmde(m, t_span=[0, 1], y0=g["disc_Y"], method="not_a_method", max_dt=0.1)
    with pytest.raises(ValueError):
        qd.solve_lmde(m, t_span=[0, 1], y0=g["disc_Y"], method="RK4", max_dt=0.1, t_eval=[0.5, 1.5])
    with pytest.raises(qd.QiskitError):
        s.solve(t_span=[0, 0.1], y0=np.ones(7), signals=sig_lists[0], method="RK4", max_dt=1e-3)


def test_batched_solver_with_sampled_pulses(qd):
    """cfg5-like: a list of simulations driven by DiscreteSignal pulses (per-simulation amplitude and width) runs
    as one sweep launch with the signal table built on the device (row f3), and matches the sequential solves
    (host signal evaluation, shared-signal kernel) -- the reference semantics results[i] == i-th individual solve
    (test_solver_classes.py:1388-1599)."""
    rng = np.random.default_rng(9)
    n, K, B, dt = 9, 2, 12, 1 / 4.5
    H0 = np.diag(np.arange(n) * 0.8) + 0.05 * (np.eye(n, k=1) + np.eye(n, k=-1))
    Hs = []
    for _ in range(K):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        Hs.append((A + A.conj().T) / (2 * np.sqrt(n)))
    y0 = np.zeros(n, dtype=complex)
    y0[0] = 1.0
    nsamp = 40

    def pulse(amp, width):
        t = (np.arange(nsamp) + 0.5) * dt
        return amp * np.exp(-0.5 * ((t - nsamp * dt / 2) / width) ** 2)

    sig_lists = [[qd.DiscreteSignal(dt, pulse(0.2 + 0.05 * b, 1.0 + 0.1 * b), carrier_freq=0.8, phase=0.0),
                  qd.DiscreteSignal(dt, 1j * pulse(0.1, 2.0), carrier_freq=1.6 + 0.01 * b)] for b in range(B)]
    s = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, rotating_frame=H0)
    T = nsamp * dt
    before = qd._abi.launch_count()
    res = s.solve(t_span=[0, T], y0=y0, signals=sig_lists, method="RK4", max_dt=dt / 2)
    launches = qd._abi.launch_count() - before
    assert isinstance(res, list) and len(res) == B and launches <= 8
    for b in (0, 5, B - 1):
        one = s.solve(t_span=[0, T], y0=y0, signals=sig_lists[b], method="RK4", max_dt=dt / 2)
        close(res[b].y[-1], npy(one.y[-1]), 1e-11)
    # chunked sweep table (tiny budget) gives the same answer
    from qiskit_dynamics_b200.solvers import fixed_step
    from qiskit_dynamics_b200.signals import compile_signal_program
    model = s.model
    model.signals = sig_lists[0]
    lists = []
    for sl in sig_lists:
        model.signals = sl
        lists.append(model.signals)
    prog = compile_signal_program(lists)
    Y0 = qd.asarray(np.repeat(y0[:, None], B, axis=1))
    yfb = model.rotating_frame.state_into_frame_basis(Y0)
    r1 = fixed_step.rk4_model_solve(model, [0, T], yfb, dt / 2, column_coefficients=lambda t: prog.table(t, yfb.device))
    r2 = fixed_step.rk4_model_solve(model, [0, T], yfb, dt / 2, column_coefficients=lambda t: prog.table(t, yfb.device),
                                    sweep_table_bytes=7 * 2 * K * B * 8)
    assert torch.equal(r1.y, r2.y)
    model.signals = None


def test_final_state_measurement(qd):
    """Row f4 on the device against fixtures made with the reference's backend_utils functions."""
    g = load_golden("measurement")
    tf = float(g["tf"])
    for name, dims, meas, slots, nslots, max_level in MEASUREMENT_CASES:
        model = qd.HamiltonianModel(static_operator=g[f"{name}_H0"], operators=[g[f"{name}_Hd"]], signals=[qd.Signal(0.1, 4.8)],
                                    rotating_frame=g[f"{name}_H0"])
        close(qd.measurement.get_lab_frame_static_hamiltonian(model), g[f"{name}_lab_h"], 1e-9)
        evals, _ = qd.measurement.get_dressed_state_decomposition(g[f"{name}_lab_h"])
        close(evals, g[f"{name}_dressed_evals"], 1e-10)
        meas_obj = qd.FinalStateMeasurement(model, dims, meas, slots, nslots, max_level, normalize_states=True)
        labels = [str(x) for x in g[f"{name}_labels"]]
        P = npy(meas_obj.probabilities(tf, g[f"{name}_Y"]))
        assert P.shape == (len(meas_obj.labels), g[f"{name}_Y"].shape[1])
        for lab, row in zip(meas_obj.labels, P):  # outcomes the fixtures never saw must carry no probability
            want = g[f"{name}_probs"][labels.index(lab)] if lab in labels else np.zeros_like(row)
            close(row, want, 1e-12)
        close(P.sum(axis=0), np.ones(P.shape[1]), 1e-13)
        # single vector in, dictionaries and counts out
        d0 = meas_obj.probabilities_dicts(tf, g[f"{name}_Y"][:, 1])[0]
        assert set(d0) <= set(meas_obj.labels) and abs(sum(d0.values()) - 1) < 1e-13
        counts = meas_obj.sample_counts(tf, g[f"{name}_Y"][:, 1], shots=200, seed=3)[0]
        assert sum(counts.values()) == 200 and set(counts) <= set(d0)
    # wide outcome table (register-free kernel path) and no normalisation
    rng = np.random.default_rng(1)
    dims = [2, 2, 2, 2, 2]
    n = 32
    H0 = np.diag(np.arange(n) * 1.0)
    model = qd.HamiltonianModel(static_operator=H0, operators=[np.eye(n)], signals=[1.0])
    m = qd.FinalStateMeasurement(model, dims, [0, 1, 2, 3, 4], max_outcome_level=None, normalize_states=False)
    Y = rng.standard_normal((n, 7)) + 1j * rng.standard_normal((n, 7))
    P = npy(m.probabilities(0.3, Y))
    assert P.shape == (32, 7)
    order = [int(lab, 2) for lab in m.labels]
    close(P, (np.abs(Y) ** 2)[order], 1e-12)


def test_lindblad(qd):
    g = load_golden("lindblad")
    H0, Hs, Lstat, Ldyn, Y = g["s_H0"], g["s_Hs"], g["s_Lstat"], g["s_Ldyn"], g["s_Y"]
    n, B = 3, Y.shape[1]
    rho = np.array([Y[:, b].reshape(n, n, order="F") for b in range(B)])
    hc, dc = g["s_hc"], g["s_dc"]
    lc = qd.LindbladCollection(static_hamiltonian=H0, hamiltonian_operators=Hs, static_dissipators=Lstat, dissipator_operators=Ldyn)
    close(lc.evaluate_rhs(hc, dc, rho), g["s_coll_rhs"])
    close(lc.evaluate_rhs(hc, dc, rho[0]), g["s_coll_rhs"][0])
    close(qd.LindbladCollection(static_hamiltonian=H0, hamiltonian_operators=Hs).evaluate_rhs(hc, None, rho), g["s_coll_rhs_hamonly"])
    close(qd.LindbladCollection(static_hamiltonian=H0, static_dissipators=Lstat).evaluate_rhs(None, None, rho), g["s_coll_rhs_statdis"])
    with pytest.raises(ValueError):
        lc.evaluate(hc, dc)
    vc = qd.VectorizedLindbladCollection(static_hamiltonian=H0, hamiltonian_operators=Hs, static_dissipators=Lstat, dissipator_operators=Ldyn)
    close(vc.evaluate(hc, dc), g["s_vcoll_eval"])
    close(vc.evaluate_rhs(hc, dc, Y), g["s_vcoll_rhs"])
    close(vc.evaluate_hamiltonian(hc), H0 + np.tensordot(hc, Hs, axes=1))
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        kw = dict(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(qd, g["s_sig"]),
                  static_dissipators=Lstat, dissipator_operators=Ldyn, dissipator_signals=sigs(qd, g["s_dsig"]),
                  rotating_frame=frame)
        mv = qd.LindbladModel(vectorized=True, **kw)
        mm = qd.LindbladModel(vectorized=False, **kw)
        assert mv.dim == 3 and mv.vectorized and not mm.vectorized
        t = 0.41
        close(mv(t), g[f"s_vec_gen_{frame_name}_fb0"], 1e-9)
        close(mv(t, Y), g[f"s_vec_rhs_{frame_name}_fb0"], 1e-9)
        close(mm(t, rho), g[f"s_mat_rhs_{frame_name}_fb0"], 1e-9)
        close(mm(t, rho[0]), g[f"s_mat_rhs1_{frame_name}_fb0"], 1e-9)
        if frame_name != "full":
            mv.in_frame_basis = True
            mm.in_frame_basis = True
            close(mv(t), g[f"s_vec_gen_{frame_name}_fb1"])
            close(mv(t, Y), g[f"s_vec_rhs_{frame_name}_fb1"])
            close(mm(t, rho), g[f"s_mat_rhs_{frame_name}_fb1"])
            oc = mv._operator_collection._operator_collection
            close(oc.static_operator, g[f"s_super_static_{frame_name}"])
            close(oc.operators, g[f"s_super_ops_{frame_name}"])
            mv.in_frame_basis = False
            mm.in_frame_basis = False
        with pytest.raises(NotImplementedError):
            mm(t)
        close(qd.solve_lmde(mv, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.05).y[-1], g[f"s_expm_{frame_name}"], 1e-9)
        close(qd.solve_lmde(mv, t_span=[0, 0.5], y0=Y, method="RK4", max_dt=0.01).y[-1], g[f"s_rk4_{frame_name}"], 1e-9)
        close(qd.solve_lmde(mv, t_span=[0, 0.5], y0=Y[:, 1], method="RK4", max_dt=0.01).y[-1], g[f"s_rk4_{frame_name}"][:, 1], 1e-9)
        close(qd.solve_lmde(mm, t_span=[0, 0.5], y0=rho, method="RK4", max_dt=0.01).y[-1], g[f"s_rk4_mat_{frame_name}"], 1e-9)
        with pytest.raises(qd.QiskitError):
            qd.solve_lmde(mm, t_span=[0, 0.5], y0=rho, method="scipy_expm", max_dt=0.05)
    # Solver with dissipators builds a LindbladModel; from_hamiltonian keeps frame + signals
    s = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, static_dissipators=Lstat, vectorized=True, rotating_frame=np.diag(H0).real)
    assert isinstance(s.model, qd.LindbladModel)
    hm = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, g["s_sig"]))
    lm = qd.LindbladModel.from_hamiltonian(hm, static_dissipators=Lstat, dissipator_operators=Ldyn,
                                           dissipator_signals=sigs(qd, g["s_dsig"]), vectorized=True)
    close(lm(0.41, Y), g["s_vec_rhs_none_fb0"], 1e-9)
    # cfg3-like: n = 27 (729), expm stepper, 1-d frame
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(27, 3, 6, 4, 2003)
    mv = qd.LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(qd, sig),
                          static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=True)
    r = qd.solve_lmde(mv, t_span=[0, 0.03], y0=Y, method="scipy_expm", max_dt=1e-2)
    assert max_col_l2(npy(r.y[-1]), g["cfg3_expm_y"]) < TOL
    close(mv(0.013, Y), g["cfg3_rhs"])


def test_magnus_orders_2_and_3(qd):
    """scipy_expm at Magnus orders 2 and 3 (fixed_step_solvers.py:80-108, 327-401): model generators run
    qdb_magnus_steps_c128, callables run qdb_magnus_terms_c128 + qdb_expm_c128; fixtures from the reference."""
    g = load_golden("magnus")
    H0, Hs, Y, sig = orc.synthetic_schrodinger(5, 2, 3, 11)
    for order in (2, 3):
        kw = dict(method="scipy_expm", magnus_order=order)
        m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig))
        close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y, max_dt=0.05, **kw).y[-1], g[f"h_noframe_o{order}"])
        m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig), rotating_frame=H0)
        close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y, max_dt=0.05, **kw).y[-1], g[f"h_full_o{order}"])
        close(qd.solve_lmde(m, t_span=[0, 0.5], y0=Y[:, 0], max_dt=0.05, **kw).y[-1], g[f"h_full_vec_o{order}"])
        r = qd.solve_lmde(m, t_span=[0.5, 0.0], y0=Y, max_dt=0.04, t_eval=[0.5, 0.31, 0.1], **kw)
        close(r.y, g[f"h_full_teval_back_o{order}"])
        A, Bm = qd.asarray(-1j * H0), qd.asarray(-1j * Hs[0])
        r = qd.solve_lmde(lambda t: A * np.cos(t) + Bm * np.sin(2 * t), t_span=[0, 1.0], y0=np.eye(5, dtype=complex),
                          max_dt=0.1, **kw)
        close(r.y[-1], g[f"callable_o{order}"])
    H0, Hs, Y, sig = orc.synthetic_schrodinger(17, 3, 6, 41)
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig), rotating_frame=H0)
    for order in (2, 3):
        r = qd.solve_lmde(m, t_span=[0, 0.3], y0=Y, method="scipy_expm", max_dt=0.03, magnus_order=order)
        assert max_col_l2(npy(r.y[-1]), g[f"h17_full_o{order}"]) < TOL
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(3, 2, 4, 4, 31)
    Lstat, Ldyn = Ls[:2], Ls[2:] + 0.02j * Ls[:2]
    dsig = [(0.3, 0.0, 0.0), (0.2, 0.11, 0.4)]
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        mv = qd.LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(qd, sig),
                              static_dissipators=Lstat, dissipator_operators=Ldyn, dissipator_signals=sigs(qd, dsig),
                              rotating_frame=frame, vectorized=True)
        for order in (2, 3):
            r = qd.solve_lmde(mv, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.05, magnus_order=order)
            close(r.y[-1], g[f"l_{frame_name}_o{order}"], 1e-9)
    with pytest.raises(qd.QiskitError):
        qd.solve_lmde(mv, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.05, magnus_order=4)
    # the C-ABI entry with room for one step at a time (step-by-step route) and for chunks of 3 steps (batched exponents,
    # commutators and exponentials) gives the same states as the full-size workspace used above
    abi = qd._abi
    from qiskit_dynamics_b200.solvers.fixed_step import expm_squarings, magnus_nodes
    from qiskit_dynamics_b200.arrays import asreal
    mv.in_frame_basis = True
    coll, n2, S, h = mv._collection(), 9, 10, 0.05
    for order in (2, 3):
        times = (h * np.arange(S))[:, None] + magnus_nodes(order)[None, :] * h
        table = mv._signal_table(times.reshape(-1))
        sq = expm_squarings(mv, table, h, order)
        outs = []
        for steps_of_room in (S, 1, 3):
            nbytes = abi.workspace_bytes(abi.WS_MAGNUS, n2, coll.num_operators, Y.shape[1], steps_of_room) + 3 * S * 8 + 256
            y = qd.asarray(Y).clone()
            abi.magnus_steps(n2, coll.operators, coll.static_operator, asreal(table, y.device), mv._frame_freqs(), times, sq, h, y, S,
                             order, workspace=torch.empty(nbytes, dtype=torch.uint8, device=y.device))
            outs.append(npy(y))
        assert max_col_l2(outs[1], outs[0]) < 1e-13 and max_col_l2(outs[2], outs[0]) < 1e-13
    mv.in_frame_basis = False


def test_nccl_sharded_sweep(qd):
    """Row (e) for the sweep path: a list of simulations split over 2 GPUs (one process each, NCCL), one gather of the
    final observables, equal to the single-GPU sweep.  Needs two devices (the 1-GPU box skips it; log of a 2-GPU run:
    profiles/r01_r_nccl_sharded_sweep.log)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29641", os.path.join(root, "tests", "_nccl_worker.py")],
                         cwd=root, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "NCCL_OK rank=0" in res.stdout and "NCCL_OK rank=1" in res.stdout


def test_time_parallel_methods(qd):
    """jax_RK4_parallel / jax_expm_parallel (fixed_step_solvers.py:206-244, 279-311, 524-613): step propagators built side
    by side and multiplied.  Against the oracle's NumPy restatement of the parallel template, and against the sequential
    solvers (same propagators; associativity of the product) -- the reference itself runs these on JAX only."""
    H0, Hs, Y, sig = orc.synthetic_schrodinger(17, 3, 6, 41)
    sp = [orc.SigSpec(*s) for s in sig]
    for frame in (None, H0, np.diag(H0).real):
        m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig), rotating_frame=frame)
        r = qd.solve_lmde(m, t_span=[0, 0.3], y0=Y, method="jax_RK4_parallel", max_dt=0.01)
        _, yo = orc.solve_hamiltonian_parallel(H0, Hs, sp, frame, [0, 0.3], Y, 0.01, kind="RK4")
        assert tuple(r.y.shape) == yo.shape and max_col_l2(npy(r.y[-1]), yo[-1]) < TOL
        seq = qd.solve_lmde(m, t_span=[0, 0.3], y0=Y, method="RK4", max_dt=0.01)
        assert max_col_l2(npy(r.y[-1]), npy(seq.y[-1])) < TOL
        for order in (1, 2, 3):
            te = [0.0, 0.11, 0.3] if order == 2 else None
            r = qd.solve_lmde(m, t_span=[0, 0.3], y0=Y, method="jax_expm_parallel", max_dt=0.03, magnus_order=order, t_eval=te)
            to, yo = orc.solve_hamiltonian_parallel(H0, Hs, sp, frame, [0, 0.3], Y, 0.03, kind="expm", magnus_order=order, t_eval=te)
            assert np.array_equal(np.asarray(r.t), np.asarray(to))
            assert max(max_col_l2(npy(r.y[i]), yo[i]) for i in range(yo.shape[0])) < TOL
    # odd step counts, a single step, backwards, vector y0, chunked workspace (7 steps per chunk at most)
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(qd, sig), rotating_frame=H0)
    for S, kw in ((1, {}), (5, {}), (37, {}), (37, dict(workspace_bytes=1))):
        r = qd.solve_lmde(m, t_span=[0.2, 0.0], y0=Y[:, 0], method="jax_RK4_parallel", max_dt=0.2 / S * (1 + 1e-12), **kw)
        _, yo = orc.solve_hamiltonian_parallel(H0, Hs, sp, H0, [0.2, 0.0], Y[:, 0], 0.2 / S * (1 + 1e-12), kind="RK4")
        assert max_col_l2(npy(r.y[-1]), yo[-1]) < TOL, S
    # vectorised Lindblad: the parallel exponential solver equals the sequential one
    H0, Hs, Ls, Yl, sig = orc.synthetic_lindblad(3, 2, 4, 5, 31)
    mv = qd.LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(qd, sig),
                          static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=True)
    a = qd.solve_lmde(mv, t_span=[0, 0.5], y0=Yl, method="jax_expm_parallel", max_dt=0.05, magnus_order=2)
    b = qd.solve_lmde(mv, t_span=[0, 0.5], y0=Yl, method="scipy_expm", max_dt=0.05, magnus_order=2)
    close(a.y[-1], npy(b.y[-1]), 1e-10)
    with pytest.raises(qd.QiskitError):
        qd.solve_lmde(lambda t: qd.asarray(H0), t_span=[0, 1], y0=Yl, method="jax_expm_parallel", max_dt=0.1)
    with pytest.raises(qd.QiskitError):
        qd.solve_lmde(mv, t_span=[0, 1], y0=Yl, method="jax_RK4_parallel")
