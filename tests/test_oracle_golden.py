"""Pins oracle/numpy_oracle.py against the reference-generated fixtures in tests/golden/.

CPU-only.  Tolerance: the oracle restates the reference's NumPy calls one for one, so these
agree to rounding (1e-12 absolute on O(1) data; most are bit-identical).
"""
import numpy as np
import pytest

from oracle import numpy_oracle as orc
from conftest import MEASUREMENT_CASES, load_golden, max_col_l2

TOL = 1e-12


def close(a, b, tol=TOL):
    np.testing.assert_allclose(a, b, rtol=0, atol=tol)


def specs(sig):
    return [orc.SigSpec(a, nu, ph) for (a, nu, ph) in np.asarray(sig)]


def test_collection():
    g = load_golden("collection")
    close(orc.collection_evaluate(g["c_real"], g["ops"], g["stat"]), g["eval_real"])
    close(orc.collection_evaluate(g["c_cplx"], g["ops"], g["stat"]), g["eval_cplx"])
    close(orc.collection_evaluate(g["c_real"], g["ops"], None), g["eval_nostat"])
    close(orc.collection_evaluate(None, None, g["stat"]), g["eval_onlystat"])
    close(orc.collection_evaluate_rhs(g["c_real"], g["ops"], g["stat"], g["yv"]), g["rhs_v"])
    close(orc.collection_evaluate_rhs(g["c_real"], g["ops"], g["stat"], g["ym"]), g["rhs_m"])
    close(orc.collection_evaluate_rhs(g["c_cplx"], g["ops"], g["stat"], g["ym"]), g["rhs_m_cplx"])
    close(orc.collection_evaluate_rhs(g["c_real"], g["ops"], None, g["ym"]), g["rhs_m_nostat"])
    close(orc.collection_evaluate_rhs(None, None, g["stat"], g["ym"]), g["rhs_m_onlystat"])
    with pytest.raises(ValueError):
        orc.collection_evaluate(None, None, None)


def test_frame():
    g = load_golden("frame")
    d, U = orc.frame_decompose(g["H"])
    close(d, g["frame_diag"])
    # eigenvectors are defined up to phase: compare projectors
    close(U @ np.diag(d) @ U.conj().T, g["frame_basis"] @ np.diag(g["frame_diag"]) @ g["frame_basis"].conj().T, 1e-11)
    d2, _ = orc.frame_decompose(-1j * g["H"])
    close(d2, g["frame_diag_anti"])
    d1, U1 = orc.frame_decompose(g["diag1d"])
    assert U1 is None
    close(d1, g["frame_diag_1d"])
    t = float(g["t"])
    U = g["frame_basis"]  # use the golden basis from here on (phase convention)
    y, op, sop = g["y"], g["op"], g["sop"]
    close(orc.state_into_frame(d, t, y), g["into_fb"])
    close(orc.state_out_of_frame(d, t, y), g["outof_fb"])
    close(U @ orc.state_into_frame(d, t, U.conj().T @ y), g["into_full"])
    close(U @ orc.state_out_of_frame(d, t, U.conj().T @ y), g["outof_full"])
    close(orc.operator_into_frame(d, t, op), g["op_into_fb"])
    close(U @ orc.operator_into_frame(d, t, U.conj().T @ op @ U) @ U.conj().T, g["op_into_full"], 1e-11)
    close(U @ orc.operator_into_frame(d, -t, U.conj().T @ op @ U) @ U.conj().T, g["op_outof_full"], 1e-11)
    close(U @ (orc.operator_into_frame(d, t, U.conj().T @ op @ U) - np.diag(d)) @ U.conj().T, g["gen_into_full"], 1e-11)
    close(orc.vectorized_map_into_frame(d, t, sop), g["vec_into_fb"])
    VU = np.kron(U.conj(), U)
    close(VU, g["vec_basis"])
    close(VU @ orc.vectorized_map_into_frame(d, t, VU.conj().T @ sop @ VU) @ VU.conj().T, g["vec_into_full"], 1e-10)
    close(orc.state_into_frame(d1, t, y), g["into_1d"])
    close(orc.operator_into_frame(d1, t, op), g["op_into_1d"])


def test_signals():
    g = load_golden("signals")
    ts = g["ts"]
    samples = g["samples"]
    d1 = orc.SigSpec(("discrete", 0.1, samples, 0.0), 1.3, 0.2)
    d2 = orc.SigSpec(("discrete", 0.1, samples, 1.0), 0.0, 0.0)
    s1 = orc.SigSpec(0.7, 2.0, 0.4)
    s2 = orc.SigSpec(lambda t: np.exp(-t**2) * (1 + 0.5j), 0.9, -1.1)
    s3 = orc.SigSpec(1.5)
    close(np.real(orc.signal_complex_value(d1, ts)), g["d1"])
    close(orc.signal_complex_value(d1, ts), g["d1_cv"])
    close(np.real(orc.signal_complex_value(d2, ts)), g["d2"])
    close(np.real(orc.signal_complex_value(s1, ts)), g["s1"])
    close(orc.signal_complex_value(s2, ts), g["s2_cv"])
    close(np.real(orc.signal_complex_value(s3, ts)), g["s3"])
    sl = [s1, s2, s3, d1, d2, [s1, s2], orc.SigSpec(2.0)]
    close(orc.signal_list_values(sl, ts), g["siglist"])
    close(orc.signal_list_values(sl, 0.123), g["siglist_scalar"])
    # bin-edge semantics with accumulated times
    d3 = orc.SigSpec(("discrete", 1 / 4.5, g["dsamp"], 0.0), 0.4, 0.0)
    env = orc.discrete_envelope(1 / 4.5, g["dsamp"], 0.0, g["tacc"])
    assert np.array_equal(env, g["d3_env_acc"])
    close(np.real(orc.signal_complex_value(d3, g["tacc"])), g["d3_acc"])


def test_step_grid():
    g = load_golden("step_grid")
    for i in range(int(g["ncases"])):
        ev = g[f"eval{i}"] if bool(g[f"has_eval{i}"]) else None
        t, h, n = orc.fixed_step_sizes(g[f"span{i}"], ev, float(g[f"maxdt{i}"]))
        assert np.array_equal(t, g[f"t{i}"])
        assert np.array_equal(h, g[f"h{i}"])
        assert np.array_equal(n, g[f"n{i}"])
    with pytest.raises(ValueError):
        orc.merge_t_args([0, 1], [0.5, 1.5])
    with pytest.raises(ValueError):
        orc.merge_t_args([0, 1], [0.7, 0.5])
    with pytest.raises(ValueError):
        orc.merge_t_args([0, 1], [[0.5]])


def test_stage_time_grid_matches_loop():
    t0, h, n = 0.0, (1 / 4.5) / 2, 20
    grid = orc.stage_time_grid(t0, h, n)
    t = t0
    for i in range(n):
        assert grid[2 * i] == t
        assert grid[2 * i + 1] == t + 0.5 * h
        assert grid[2 * i + 2] == t + h
        t = t + h


def test_hamiltonian_model():
    g = load_golden("hamiltonian_model")
    H0, Hs, Y, sig = g["H0"], g["Hs"], g["Y"], g["sig"]
    sp = specs(sig)
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        Gd, G, d, U = orc.generator_model_operators(H0, Hs, frame)
        # eigenvector phases are arbitrary -> compare through out-of-basis quantities
        for i, t in enumerate(g["ts"]):
            yfb = Y if U is None else U.conj().T @ Y
            out = orc.model_rhs(t, yfb, sp, G, Gd, d)
            out = out if U is None else U @ out
            close(out, g[f"rhs_{frame_name}_fb0_{i}"], 1e-11)
            gen = orc.model_generator(t, sp, G, Gd, d)
            gen = gen if U is None else U @ gen @ U.conj().T
            close(gen, g[f"gen_{frame_name}_fb0_{i}"], 1e-11)
        if U is None:
            close(orc.model_rhs(0.37, Y, sp, G, Gd, d), g[f"rhs_{frame_name}_fb1_1"], 1e-11)
            close(orc.model_rhs(0.37, Y[:, 0], sp, G, Gd, d), g[f"rhsvec_{frame_name}_fb1"], 1e-11)
            close(G, g[f"ops_{frame_name}"])
            close(Gd, g[f"stat_{frame_name}"])
    Gd, G, d, U = orc.generator_model_operators(None, Hs, H0)
    close(np.diag(Gd), np.diag(g["stat_nostatic"]))
    Gd, G, d, U = orc.generator_model_operators(g["Gd"], g["Gs"], -1j * H0, hamiltonian=False)
    out = U @ orc.model_rhs(0.37, U.conj().T @ Y, sp, G, Gd, d)
    close(out, g["genmodel_rhs"], 1e-10)
    gen = U @ orc.model_generator(0.37, sp, G, Gd, d) @ U.conj().T
    close(gen, g["genmodel_gen"], 1e-10)


def test_rk4_solves():
    g = load_golden("rk4_solves")
    # cfg1
    _, ys = orc.solve_hamiltonian(g["cfg1_H0"], g["cfg1_H1"][None], [orc.SigSpec(1.0, 5.0)], g["cfg1_H0"],
                                  [0, 10.0], g["cfg1_y0"], 1e-3)
    close(ys, g["cfg1_y"], 1e-10)
    close(ys, g["cfg1_solver_y"], 1e-10)
    # cfg4-like
    H0, Hs, Y, sig = orc.synthetic_schrodinger(128, 8, 8, 2004)
    close(np.array([np.sum(np.abs(a)) for a in (H0, Hs, Y)] + [np.sum(a).real for a in (H0, Hs, Y)]), g["cfg4_check"], 1e-9)
    sp = specs(sig)
    _, ys = orc.solve_hamiltonian(H0, Hs, sp, H0, [0, 0.05], Y, 1e-3)
    close(ys[-1], g["cfg4_y"], 1e-11)
    t, ys = orc.solve_hamiltonian(H0, Hs, sp, H0, [0, 0.02], Y, 1e-3, t_eval=[0.0, 0.005, 0.0125, 0.02])
    assert np.array_equal(t, g["cfg4_teval_t"])
    close(ys, g["cfg4_teval_y"], 1e-11)
    _, ys = orc.solve_hamiltonian(H0, Hs, sp, H0, [0, 1e-3], Y, 1e-3)
    close(ys[-1], g["cfg4_onestep_y"], 1e-12)
    _, ys = orc.solve_hamiltonian(H0, Hs, sp, H0, [0.02, 0.0], Y, 1e-3)
    close(ys[-1], g["cfg4_back_y"], 1e-11)
    # cfg2-like sweep
    H0, Hs, Y, sig = orc.synthetic_schrodinger(32, 8, 1, 2002)
    B = 16
    per_col = [[orc.SigSpec(a * (0.5 + b / B), nu, ph) for (a, nu, ph) in sig] for b in range(B)]
    _, ys = orc.solve_hamiltonian_sweep(H0, Hs, per_col, H0, [0, 0.1], np.repeat(Y, B, axis=1), 1e-3)
    close(ys[-1], g["cfg2_y"], 1e-11)
    # odd dimension
    H0, Hs, Y, sp = g["odd_H0"], g["odd_Hs"], g["odd_Y"], specs(g["odd_sig"])
    close(orc.solve_hamiltonian(H0, Hs, sp, None, [0, 0.5], Y, 0.01)[1][-1], g["odd_noframe_y"], 1e-12)
    fr = np.diag(H0).real
    close(orc.solve_hamiltonian(H0, Hs, sp, fr, [0, 0.5], Y, 0.01)[1][-1], g["odd_diagframe_y"], 1e-12)
    close(orc.solve_hamiltonian(H0, Hs, sp, fr, [0, 0.5], Y[:, 0], 0.01)[1][-1], g["odd_vec_y"], 1e-12)
    close(orc.solve_hamiltonian(H0, Hs, sp, fr, [0, 0.5], np.eye(5, dtype=complex), 0.01)[1][-1], g["odd_eye_y"], 1e-12)
    close(orc.solve_hamiltonian(H0, Hs, sp, fr, [0, 0.5], Y, 0.01, method="scipy_expm")[1][-1], g["odd_expm_y"], 1e-12)
    close(orc.solve_hamiltonian(H0, Hs, sp, H0, [0, 0.5], Y, 0.01, method="scipy_expm")[1][-1], g["odd_expm_fullframe_y"], 1e-11)
    # DiscreteSignal drive, stage times on bin edges
    dt = float(g["disc_dt"])
    dsp = [orc.SigSpec(("discrete", dt, g["disc_samples"][j], 0.0), 0.3 * (j + 1), 0.1 * j) for j in range(2)]
    close(orc.solve_hamiltonian(g["disc_H0"], g["disc_Hs"], dsp, g["disc_H0"], [0, 2.0], g["disc_Y"], dt / 2)[1][-1],
          g["disc_y"], 1e-11)
    close(orc.solve_hamiltonian(g["disc_H0"], g["disc_Hs"], dsp, None, [0, 2.0], g["disc_Y"], dt / 2)[1][-1],
          g["disc_noframe_y"], 1e-11)


def test_lindblad():
    g = load_golden("lindblad")
    H0, Hs, Lstat, Ldyn, Y = g["s_H0"], g["s_Hs"], g["s_Lstat"], g["s_Ldyn"], g["s_Y"]
    sp, dsp = specs(g["s_sig"]), specs(g["s_dsig"])
    n, B = 3, Y.shape[1]
    rho = np.array([Y[:, b].reshape(n, n, order="F") for b in range(B)])
    hc, dc = g["s_hc"], g["s_dc"]
    # collection-level
    close(orc.lindblad_rhs_matrix(hc, dc, H0, Hs, Lstat, Ldyn, rho), g["s_coll_rhs"])
    close(orc.lindblad_rhs_matrix(hc, None, H0, Hs, None, None, rho), g["s_coll_rhs_hamonly"])
    close(orc.lindblad_rhs_matrix(None, None, H0, None, Lstat, None, rho), g["s_coll_rhs_statdis"])
    S, ops = orc.vectorized_lindblad_collection(H0, Hs, Lstat, Ldyn)
    c = np.append(hc, dc)
    close(orc.collection_evaluate(c, ops, S), g["s_vcoll_eval"])
    close(orc.collection_evaluate_rhs(c, ops, S, Y), g["s_vcoll_rhs"])
    # vec == mat consistency of the oracle itself (reference test_operator_collections.py:550-706)
    vec_of_mat = np.stack([m.flatten(order="F") for m in orc.lindblad_rhs_matrix(hc, dc, H0, Hs, Lstat, Ldyn, rho)], axis=-1)
    close(vec_of_mat, g["s_vcoll_rhs"])
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        for method, key, mdt in (("scipy_expm", "s_expm_", 0.05), ("RK4", "s_rk4_", 0.01)):
            _, ys = orc.solve_vectorized_lindblad(H0, Hs, sp, Lstat, Ldyn, dsp, frame, [0, 0.5], Y, mdt, method)
            close(ys[-1], g[key + frame_name], 1e-11)
        # vectorised RK4 == non-vectorised RK4 (reference checked in SURVEY f2)
        close(np.stack([m.flatten(order="F") for m in g[f"s_rk4_mat_{frame_name}"]], axis=-1), g["s_rk4_" + frame_name], 1e-11)
        if frame is None or np.asarray(frame).ndim == 1:
            Hd_, Hops_, Ds_, Do_, d, U = orc.lindblad_model_operators(H0, Hs, Lstat, Ldyn, frame)
            S, ops = orc.vectorized_lindblad_collection(Hd_, Hops_, Ds_, Do_)
            close(S, g[f"s_super_static_{frame_name}"])
            close(ops, g[f"s_super_ops_{frame_name}"])
            # frame-phase identity used by the CUDA path: mu_a = lam_i - lam_k
            if d is not None:
                t = 0.41
                cc = np.append(orc.signal_list_values(sp, t), orc.signal_list_values(dsp, t))
                mu = orc.vec_frame_phase(d)
                p = np.exp(-1j * mu * t)
                rhs = p.conj()[:, None] * (orc.collection_evaluate(cc, ops, S) @ (p[:, None] * Y))
                close(rhs, g[f"s_vec_rhs_{frame_name}_fb1"], 1e-12)
                gen = orc.collection_evaluate(cc, ops, S) * np.outer(p.conj(), p)
                close(gen, g[f"s_vec_gen_{frame_name}_fb1"], 1e-12)
    # cfg3-like
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(27, 3, 6, 4, 2003)
    close(np.array([np.sum(np.abs(a)) for a in (H0, Hs, Ls, Y)] + [np.sum(a).real for a in (H0, Hs, Ls, Y)]), g["cfg3_check"], 1e-9)
    _, ys = orc.solve_vectorized_lindblad(H0, Hs, specs(sig), Ls, None, None, np.diag(H0).real, [0, 0.03], Y, 1e-2)
    close(ys[-1], g["cfg3_expm_y"], 1e-12)


def test_measurement_post_processing():
    """Row f4: dressed-state sorting, out-of-frame, normalisation and memory-slot probabilities against fixtures
    produced by the reference's backend_utils functions."""
    g = load_golden("measurement")
    tf = float(g["tf"])
    for name, dims, meas, slots, nslots, max_level in MEASUREMENT_CASES:
        evals, dressed = orc.dressed_state_decomposition(g[f"{name}_lab_h"])
        close(evals, g[f"{name}_dressed_evals"], 1e-10)
        overlap = np.abs(np.sum(dressed.conj() * g[f"{name}_dressed_states"], axis=0))  # eigenvector phases are free
        close(overlap, np.ones_like(overlap), 1e-10)
        labels = [str(x) for x in g[f"{name}_labels"]]
        Y = g[f"{name}_Y"]
        for b in range(Y.shape[1]):
            d = orc.final_state_memory_probabilities(Y[:, b], tf, g[f"{name}_H0"], g[f"{name}_dressed_states"], dims, meas,
                                                     slots, nslots, max_level, normalize=True)
            assert set(d) <= set(labels)
            close(np.array([d.get(lab, 0.0) for lab in labels]), g[f"{name}_probs"][:, b])
    # the reference's own example of the slot mapping (computed with backend_utils._get_memory_slot_probabilities)
    assert orc.memory_slot_probabilities({"00": 0.1, "01": 0.2, "10": 0.3, "12": 0.4}, [0, 2], 3, 1) == \
        {"000": 0.1, "001": 0.2, "100": 0.3, "101": 0.4}


def test_magnus_orders():
    """Row a9 beyond first order: scipy_expm_solver(magnus_order=2, 3) of the reference
    (solvers/fixed_step_solvers.py:327-401) against the oracle's restatement."""
    g = load_golden("magnus")
    H0, Hs, Y, sig = orc.synthetic_schrodinger(5, 2, 3, 11)
    close(np.array([np.sum(np.abs(a)) for a in (H0, Hs, Y)] + [np.sum(a).real for a in (H0, Hs, Y)]), g["h_check"], 1e-9)
    sp = [orc.SigSpec(*s) for s in sig]
    for order in (2, 3):
        kw = dict(method="scipy_expm", magnus_order=order)
        close(orc.solve_hamiltonian(H0, Hs, sp, None, [0, 0.5], Y, 0.05, **kw)[1][-1], g[f"h_noframe_o{order}"])
        close(orc.solve_hamiltonian(H0, Hs, sp, H0, [0, 0.5], Y, 0.05, **kw)[1][-1], g[f"h_full_o{order}"], 1e-11)
        close(orc.solve_hamiltonian(H0, Hs, sp, H0, [0, 0.5], Y[:, 0], 0.05, **kw)[1][-1], g[f"h_full_vec_o{order}"], 1e-11)
        close(orc.solve_hamiltonian(H0, Hs, sp, H0, [0.5, 0.0], Y, 0.04, t_eval=[0.5, 0.31, 0.1], **kw)[1],
              g[f"h_full_teval_back_o{order}"], 1e-11)
        A, Bm = -1j * H0, -1j * Hs[0]
        ys = orc.fixed_step_solve(orc.magnus_step(order), lambda t: A * np.cos(t) + Bm * np.sin(2 * t), [0, 1.0],
                                  np.eye(5, dtype=complex), 0.1)[1]
        close(ys[-1], g[f"callable_o{order}"])
    # the orders differ from each other by far more than the tolerance (the fixtures do discriminate)
    assert np.max(np.abs(g["h_full_o2"] - g["h_full_o3"])) > 1e-9
    H0, Hs, Y, sig = orc.synthetic_schrodinger(17, 3, 6, 41)
    sp = [orc.SigSpec(*s) for s in sig]
    for order in (2, 3):
        close(orc.solve_hamiltonian(H0, Hs, sp, H0, [0, 0.3], Y, 0.03, method="scipy_expm", magnus_order=order)[1][-1],
              g[f"h17_full_o{order}"], 1e-11)
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(3, 2, 4, 4, 31)
    Lstat, Ldyn = Ls[:2], Ls[2:] + 0.02j * Ls[:2]
    sp = [orc.SigSpec(*s) for s in sig]
    dsp = [orc.SigSpec(0.3, 0.0, 0.0), orc.SigSpec(0.2, 0.11, 0.4)]
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        for order in (2, 3):
            _, ys = orc.solve_vectorized_lindblad(H0, Hs, sp, Lstat, Ldyn, dsp, frame, [0, 0.5], Y, 0.05, "scipy_expm",
                                                  magnus_order=order)
            close(ys[-1], g[f"l_{frame_name}_o{order}"], 1e-11)


def test_parallel_template_agrees_with_sequential():
    """The NumPy restatement of the reference's JAX-only parallel template (fixed_step_solvers.py:206-244, 524-613) has no
    fixture of its own: it must reproduce the sequential solvers, which are pinned above, from the same step propagators
    (RK4 on a linear system IS its propagator; exponential steps at every Magnus order), and the golden RK4 / expm solves."""
    g = load_golden("rk4_solves")
    H0, Hs, Y, sp = g["odd_H0"], g["odd_Hs"], g["odd_Y"], specs(g["odd_sig"])
    fr = np.diag(H0).real
    close(orc.solve_hamiltonian_parallel(H0, Hs, sp, None, [0, 0.5], Y, 0.01, kind="RK4")[1][-1], g["odd_noframe_y"], 1e-12)
    close(orc.solve_hamiltonian_parallel(H0, Hs, sp, fr, [0, 0.5], Y, 0.01, kind="RK4")[1][-1], g["odd_diagframe_y"], 1e-12)
    close(orc.solve_hamiltonian_parallel(H0, Hs, sp, fr, [0, 0.5], Y, 0.01, kind="expm")[1][-1], g["odd_expm_y"], 1e-12)
    close(orc.solve_hamiltonian_parallel(H0, Hs, sp, H0, [0, 0.5], Y, 0.01, kind="expm")[1][-1], g["odd_expm_fullframe_y"], 1e-11)
    gm = load_golden("magnus")
    H0, Hs, Y, sig = orc.synthetic_schrodinger(5, 2, 3, 11)
    sp = [orc.SigSpec(*s) for s in sig]
    for order in (2, 3):
        close(orc.solve_hamiltonian_parallel(H0, Hs, sp, H0, [0, 0.5], Y, 0.05, kind="expm", magnus_order=order)[1][-1],
              gm[f"h_full_o{order}"], 1e-11)
        t, ys = orc.solve_hamiltonian_parallel(H0, Hs, sp, H0, [0.5, 0.0], Y, 0.04, kind="expm", magnus_order=order,
                                               t_eval=[0.5, 0.31, 0.1])
        close(ys, gm[f"h_full_teval_back_o{order}"], 1e-11)


# ---------------------------------------------------------------------------------------------
# full-size configurations: the oracle on the chosen columns == the unmodified reference on the same columns
# (tests/golden/fullsize.npz); the GPU side of the same comparison is tests/test_fullsize_gpu.py
# ---------------------------------------------------------------------------------------------


def test_fullsize_columns_oracle_equals_reference():
    import bench_workloads as W
    import fullsize_cases as fc
    g = load_golden("fullsize")
    assert np.array_equal(g["cfg4_cols"], W.parity_columns(4096)) and g["cfg4_cols"].size == 32
    assert max_col_l2(fc.oracle_cfg4(g["cfg4_cols"]), g["cfg4_y"]) < 1e-12       # 1000 RK4 steps, n = 128
    assert max_col_l2(fc.oracle_cfg4(g["cfg4_ragged_cols"], t_end=0.1), g["cfg4_ragged_y"]) < 1e-12
    assert max_col_l2(fc.oracle_cfg2(g["cfg2_cols"][:8]), g["cfg2_y"][:, :8]) < 1e-12  # 1000 steps each, per-column signals
    assert max_col_l2(fc.oracle_cfg3(g["cfg3_cols"][:4]), g["cfg3_y"][:, :4]) < 1e-12  # 20 exponentials of 729 x 729
    assert max_col_l2(fc.oracle_cfg5(g["cfg5_cols"]), g["cfg5_y"]) < 1e-12       # bin-edge stage times, 64 steps
    # memory-slot probabilities of the reference's post-processing chain
    H0, ops, freqs = W.cfg5_system()
    dims, msub, mslots = W.cfg5_measurement()
    _, dressed = orc.dressed_state_decomposition(H0)
    labels = [str(x) for x in g["cfg5_labels"]]
    for i in (0, 7, 31):
        pd = orc.final_state_memory_probabilities(g["cfg5_y"][:, i], fc.CFG5_NSAMP * W.CFG5_DT, H0, dressed, dims, msub, mslots,
                                                  max_outcome_value=1)
        close(np.array([pd.get(lab, 0.0) for lab in labels]), g["cfg5_probs"][:, i])
