"""Generate golden fixtures by running the UNMODIFIED reference (NumPy path).

Run in the dev container where /root/reference is mounted:

    python tests/golden/make_golden.py

Writes tests/golden/*.npz (inputs that are small + reference outputs).  Large inputs are
regenerated from seeds by `oracle.numpy_oracle.synthetic_*`; a checksum of them is stored so a
change in NumPy's generator stream would be detected.  Reference: qiskit-dynamics 0.6.0
(/root/reference/qiskit_dynamics/VERSION.txt), imported through oracle/ref_shim.py.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle.ref_shim as ref_shim  # noqa: E402

if not ref_shim.reference_available():
    raise SystemExit("reference not available (baseline/_ref or /root/reference); cannot regenerate goldens")

from qiskit_dynamics import Signal, DiscreteSignal, solve_lmde, Solver  # noqa: E402
from qiskit_dynamics.signals import SignalList, SignalSum  # noqa: E402
from qiskit_dynamics.models import (  # noqa: E402
    HamiltonianModel, LindbladModel, GeneratorModel, RotatingFrame)
from qiskit_dynamics.models.operator_collections import (  # noqa: E402
    OperatorCollection, LindbladCollection, VectorizedLindbladCollection)
from qiskit_dynamics.solvers.fixed_step_solvers import get_fixed_step_sizes  # noqa: E402

from oracle import numpy_oracle as orc  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def checksum(*arrays):
    return np.array([np.sum(np.abs(a)) for a in arrays] + [np.sum(a).real for a in arrays])


def sigs(spec):
    return [Signal(a, nu, ph) for (a, nu, ph) in spec]


# ---------------------------------------------------------------------------------------------
def gen_collection():
    rng = np.random.default_rng(342)
    K, n, B = 5, 6, 7
    ops = rng.standard_normal((K, n, n)) + 1j * rng.standard_normal((K, n, n))
    stat = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    c_real = rng.standard_normal(K)
    c_cplx = rng.standard_normal(K) + 1j * rng.standard_normal(K)
    yv = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ym = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    full = OperatorCollection(static_operator=stat, operators=ops)
    nostat = OperatorCollection(operators=ops)
    onlystat = OperatorCollection(static_operator=stat)
    save("collection", ops=ops, stat=stat, c_real=c_real, c_cplx=c_cplx, yv=yv, ym=ym,
         eval_real=full.evaluate(c_real), eval_cplx=full.evaluate(c_cplx),
         eval_nostat=nostat.evaluate(c_real), eval_onlystat=onlystat.evaluate(None),
         rhs_v=full.evaluate_rhs(c_real, yv), rhs_m=full.evaluate_rhs(c_real, ym),
         rhs_m_cplx=full.evaluate_rhs(c_cplx, ym), rhs_m_nostat=nostat(c_real, ym),
         rhs_m_onlystat=onlystat(None, ym))


def gen_frame():
    rng = np.random.default_rng(4531)
    n, B = 5, 3
    H = orc.herm(rng, n) * 3
    rf = RotatingFrame(H)
    rf_anti = RotatingFrame(-1j * H)
    rf1d = RotatingFrame(np.array([1.0, -0.5, 2.0, 0.25, 3.0]))
    y = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    op = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    sop = rng.standard_normal((n * n, n * n)) + 1j * rng.standard_normal((n * n, n * n))
    t = 0.7312
    save("frame", H=H, y=y, op=op, sop=sop, t=np.array(t), diag1d=np.array([1.0, -0.5, 2.0, 0.25, 3.0]),
         frame_diag=rf.frame_diag, frame_basis=rf.frame_basis,
         frame_diag_anti=rf_anti.frame_diag,
         frame_diag_1d=rf1d.frame_diag,
         into_fb=rf.state_into_frame(t, y, y_in_frame_basis=True, return_in_frame_basis=True),
         outof_fb=rf.state_out_of_frame(t, y, y_in_frame_basis=True, return_in_frame_basis=True),
         into_full=rf.state_into_frame(t, y),
         outof_full=rf.state_out_of_frame(t, y),
         op_into_fb=rf.operator_into_frame(t, op, operator_in_frame_basis=True, return_in_frame_basis=True),
         op_into_full=rf.operator_into_frame(t, op),
         op_outof_full=rf.operator_out_of_frame(t, op),
         gen_into_full=rf.generator_into_frame(t, op),
         gen_outof_full=rf.generator_out_of_frame(t, op),
         gen_into_fb0=rf.generator_into_frame(0.0, op, return_in_frame_basis=True),
         vec_into_fb=rf.vectorized_map_into_frame(t, sop, operator_in_frame_basis=True, return_in_frame_basis=True),
         vec_into_full=rf.vectorized_map_into_frame(t, sop),
         vec_basis=rf.vectorized_frame_basis,
         state_into_basis=rf.state_into_frame_basis(y), state_outof_basis=rf.state_out_of_frame_basis(y),
         op_into_basis=rf.operator_into_frame_basis(op), op_outof_basis=rf.operator_out_of_frame_basis(op),
         into_1d=rf1d.state_into_frame(t, y), op_into_1d=rf1d.operator_into_frame(t, op))


def gen_signals():
    ts = np.array([0.0, 0.05, 0.1, 0.1 + 1e-17, 0.2, 0.29999999999999993, 0.3, 0.30000000000000004,
                   0.31, -0.01, 1.0, 1.2, 1.5555555555555556, 2.9, 3.3])
    samples = np.array([1.0, 2.0, 3.0 + 1j])
    d1 = DiscreteSignal(dt=0.1, samples=samples, carrier_freq=1.3, phase=0.2)
    d2 = DiscreteSignal(dt=0.1, samples=samples, start_time=1.0, carrier_freq=0.0)
    s1 = Signal(0.7, 2.0, 0.4)
    s2 = Signal(lambda t: np.exp(-t**2) * (1 + 0.5j), carrier_freq=0.9, phase=-1.1)
    s3 = Signal(1.5)
    ssum = s1 + s2
    sprod = s1 * s2
    dprod = d1 * d1
    sl = SignalList([s1, s2, s3, d1, d2, ssum, 2.0])
    # accumulated stage times with dt = 1/4.5, h = dt/2 (bin-edge trap)
    dt = 1 / 4.5
    h = dt / 2
    tacc = [0.0]
    for _ in range(20):
        tacc.append(tacc[-1] + h)
    tacc = np.array(tacc)
    rng = np.random.default_rng(7)
    dsamp = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    d3 = DiscreteSignal(dt=dt, samples=dsamp, carrier_freq=0.4)
    save("signals", ts=ts, samples=samples,
         d1=d1(ts), d1_cv=d1.complex_value(ts), d2=d2(ts), s1=s1(ts), s2=s2(ts), s2_cv=s2.complex_value(ts),
         s3=s3(ts), ssum=ssum(ts), sprod=sprod(ts), dprod=dprod(ts),
         siglist=sl(ts), siglist_scalar=sl(0.123), siglist_cv=sl.complex_value(ts),
         drift=sl.drift, tacc=tacc, dsamp=dsamp, d3_acc=d3(tacc), d3_env_acc=d3.envelope(tacc),
         conj_d1=d1.conjugate().complex_value(ts))


def gen_step_grid():
    cases = [([0.0, 1.0], None, 0.1), ([0.0, 1.0], None, 0.3), ([0.0, 1.0], [0.25, 0.5, 0.9], 0.1),
             ([1.0, 0.0], [0.75, 0.5], 0.2), ([0.0, 1.0], None, 1e-3), ([0.0, 0.05], None, 1e-3),
             ([0.0, 10.0], None, 1e-3), ([0.0, 1.0], [0.0, 0.5, 1.0], 0.11), ([0.0, 0.2], None, 1e-2),
             ([0.0, 1.0], None, 5.0)]
    out = {}
    for i, (span, ev, mdt) in enumerate(cases):
        t_list, h_list, n_list = get_fixed_step_sizes(span, ev, mdt)
        out[f"span{i}"] = np.array(span)
        out[f"eval{i}"] = np.array([]) if ev is None else np.array(ev)
        out[f"has_eval{i}"] = np.array(ev is not None)
        out[f"maxdt{i}"] = np.array(mdt)
        out[f"t{i}"] = np.array(t_list)
        out[f"h{i}"] = np.array(h_list)
        out[f"n{i}"] = np.array(n_list)
    out["ncases"] = np.array(len(cases))
    save("step_grid", **out)


def gen_hamiltonian_model():
    n, K, B = 8, 3, 5
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 99)
    out = dict(H0=H0, Hs=Hs, Y=Y, sig=np.array(sig), ts=np.array([0.0, 0.37, 2.5]))
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig), rotating_frame=frame)
        for fb in (False, True):
            m.in_frame_basis = fb
            for i, t in enumerate(out["ts"]):
                out[f"rhs_{frame_name}_fb{int(fb)}_{i}"] = m(t, Y)
                out[f"gen_{frame_name}_fb{int(fb)}_{i}"] = m(t)
            out[f"rhsvec_{frame_name}_fb{int(fb)}"] = m(0.37, Y[:, 0])
        m.in_frame_basis = True
        out[f"stat_{frame_name}"] = (np.zeros((n, n), complex) if m._operator_collection.static_operator is None
                                     else m._operator_collection.static_operator)
        out[f"ops_{frame_name}"] = m._operator_collection.operators
    # no static operator, frame only
    m = HamiltonianModel(operators=Hs, signals=sigs(sig), rotating_frame=H0, in_frame_basis=True)
    out["rhs_nostatic"] = m(0.37, Y)
    out["stat_nostatic"] = m._operator_collection.static_operator
    # GeneratorModel (no -i fold), non-Hermitian generator
    rng = np.random.default_rng(5)
    Gs = rng.standard_normal((K, n, n)) + 1j * rng.standard_normal((K, n, n))
    Gd = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    gm = GeneratorModel(static_operator=Gd, operators=Gs, signals=sigs(sig), rotating_frame=-1j * H0)
    out["Gs"] = Gs
    out["Gd"] = Gd
    out["genmodel_rhs"] = gm(0.37, Y)
    out["genmodel_gen"] = gm(0.37)
    save("hamiltonian_model", **out)


def gen_rk4_solves():
    out = {}
    # cfg1: 2-qubit, 1 Rabi drive, RK4, T shortened to 1.0 for fixture size
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Z = np.diag([1.0, -1.0]).astype(complex)
    I2 = np.eye(2, dtype=complex)
    H0 = 2 * np.pi * 5 * (np.kron(Z, I2) + np.kron(I2, Z)) / 2
    H1 = 2 * np.pi * 0.1 * np.kron(X, I2) / 2
    y0 = np.array([1.0, 0, 0, 0], dtype=complex)
    m = HamiltonianModel(static_operator=H0, operators=[H1], signals=[Signal(1.0, 5.0)], rotating_frame=H0)
    r = solve_lmde(m, t_span=[0, 10.0], y0=y0, method="RK4", max_dt=1e-3)
    out["cfg1_H0"], out["cfg1_H1"], out["cfg1_y0"], out["cfg1_y"] = H0, H1, y0, r.y
    # same through Solver.solve with array y0 (plumbing)
    s = Solver(static_hamiltonian=H0, hamiltonian_operators=[H1], rotating_frame=H0)
    r2 = s.solve(t_span=[0, 10.0], y0=y0, signals=[Signal(1.0, 5.0)], method="RK4", max_dt=1e-3)
    out["cfg1_solver_y"] = r2.y

    # cfg4-like: n=128, K=8, B=8 columns, 50 steps, frame = H0
    n, K, B = 128, 8, 8
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 2004)
    out["cfg4_check"] = checksum(H0, Hs, Y)
    m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig), rotating_frame=H0)
    r = solve_lmde(m, t_span=[0, 0.05], y0=Y, method="RK4", max_dt=1e-3)
    out["cfg4_y"] = r.y[-1]
    m.in_frame_basis = True
    out["cfg4_rhs_fb"] = np.array([m(t, Y) for t in (0.0, 0.0135, 0.05)])
    m.in_frame_basis = False
    # t_eval + intermediate results, and a single RK4 step
    r = solve_lmde(m, t_span=[0, 0.02], y0=Y, method="RK4", max_dt=1e-3, t_eval=[0.0, 0.005, 0.0125, 0.02])
    out["cfg4_teval_t"], out["cfg4_teval_y"] = np.array(r.t), r.y
    r = solve_lmde(m, t_span=[0, 1e-3], y0=Y, method="RK4", max_dt=1e-3)
    out["cfg4_onestep_y"] = r.y[-1]
    # backwards integration
    r = solve_lmde(m, t_span=[0.02, 0.0], y0=Y, method="RK4", max_dt=1e-3)
    out["cfg4_back_y"] = r.y[-1]

    # cfg2-like sweep: n=32, K=8, B=16 per-column amplitudes a_{b,j} = a_j (0.5 + b/B), single y0
    n, K, B = 32, 8, 16
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, 1, 2002)
    out["cfg2_check"] = checksum(H0, Hs, Y)
    m = HamiltonianModel(static_operator=H0, operators=Hs, rotating_frame=H0)
    cols = []
    for b in range(B):
        m.signals = [Signal(a * (0.5 + b / B), nu, ph) for (a, nu, ph) in sig]
        r = solve_lmde(m, t_span=[0, 0.1], y0=Y[:, 0], method="RK4", max_dt=1e-3)
        cols.append(r.y[-1])
    out["cfg2_y"] = np.stack(cols, axis=-1)

    # odd dimension (padding path): n=5, K=2, B=3, no frame / 1-d frame
    n, K, B = 5, 2, 3
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 11)
    out["odd_H0"], out["odd_Hs"], out["odd_Y"], out["odd_sig"] = H0, Hs, Y, np.array(sig)
    m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig))
    out["odd_noframe_y"] = solve_lmde(m, t_span=[0, 0.5], y0=Y, method="RK4", max_dt=0.01).y[-1]
    m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig), rotating_frame=np.diag(H0).real)
    out["odd_diagframe_y"] = solve_lmde(m, t_span=[0, 0.5], y0=Y, method="RK4", max_dt=0.01).y[-1]
    out["odd_vec_y"] = solve_lmde(m, t_span=[0, 0.5], y0=Y[:, 0], method="RK4", max_dt=0.01).y[-1]
    # square y0 (propagator) as in test_fixed_step_solvers
    out["odd_eye_y"] = solve_lmde(m, t_span=[0, 0.5], y0=np.eye(n, dtype=complex), method="RK4", max_dt=0.01).y[-1]
    # expm stepper on the Hamiltonian model
    out["odd_expm_y"] = solve_lmde(m, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.01).y[-1]
    m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig), rotating_frame=H0)
    out["odd_expm_fullframe_y"] = solve_lmde(m, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.01).y[-1]

    # DiscreteSignal drive with max_dt = sample width: every stage time on a bin edge (A.4)
    n, K, B = 6, 2, 4
    H0, Hs, Y, _ = orc.synthetic_schrodinger(n, K, B, 21)
    dt = 1 / 4.5
    rng = np.random.default_rng(22)
    samp = rng.standard_normal((K, 9)) + 1j * rng.standard_normal((K, 9))
    dsigs = [DiscreteSignal(dt=dt, samples=samp[j], carrier_freq=0.3 * (j + 1), phase=0.1 * j) for j in range(K)]
    out["disc_H0"], out["disc_Hs"], out["disc_Y"], out["disc_samples"], out["disc_dt"] = H0, Hs, Y, samp, np.array(dt)
    m = HamiltonianModel(static_operator=H0, operators=Hs, signals=dsigs, rotating_frame=H0)
    out["disc_y"] = solve_lmde(m, t_span=[0, 2.0], y0=Y, method="RK4", max_dt=dt / 2).y[-1]
    m2 = HamiltonianModel(static_operator=H0, operators=Hs, signals=dsigs)
    out["disc_noframe_y"] = solve_lmde(m2, t_span=[0, 2.0], y0=Y, method="RK4", max_dt=dt / 2).y[-1]
    save("rk4_solves", **out)


def gen_lindblad():
    out = {}
    # small: n=3, 2 ham ops, 2 static + 2 time-dependent dissipators
    n, K, B = 3, 2, 4
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 4, B, 31)
    Lstat, Ldyn = Ls[:2], Ls[2:] + 0.02j * Ls[:2]
    dsig = [(0.3, 0.0, 0.0), (0.2, 0.11, 0.4)]
    out.update(s_H0=H0, s_Hs=Hs, s_Lstat=Lstat, s_Ldyn=Ldyn, s_Y=Y, s_sig=np.array(sig), s_dsig=np.array(dsig))
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        kw = dict(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(sig),
                  static_dissipators=Lstat, dissipator_operators=Ldyn, dissipator_signals=sigs(dsig),
                  rotating_frame=frame)
        mv = LindbladModel(vectorized=True, **kw)
        mm = LindbladModel(vectorized=False, **kw)
        t = 0.41
        for fb in (False, True):
            mv.in_frame_basis = fb
            mm.in_frame_basis = fb
            out[f"s_vec_gen_{frame_name}_fb{int(fb)}"] = mv(t)
            out[f"s_vec_rhs_{frame_name}_fb{int(fb)}"] = mv(t, Y)
            rho = Y.T.reshape(B, n, n, order="F")[:, :, :]
            rho = np.array([Y[:, b].reshape(n, n, order="F") for b in range(B)])
            out[f"s_mat_rhs_{frame_name}_fb{int(fb)}"] = mm(t, rho)
            out[f"s_mat_rhs1_{frame_name}_fb{int(fb)}"] = mm(t, rho[0])
        mv.in_frame_basis = False
        mm.in_frame_basis = False
        out[f"s_expm_{frame_name}"] = solve_lmde(mv, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.05).y[-1]
        out[f"s_rk4_{frame_name}"] = solve_lmde(mv, t_span=[0, 0.5], y0=Y, method="RK4", max_dt=0.01).y[-1]
        rho = np.array([Y[:, b].reshape(n, n, order="F") for b in range(B)])
        out[f"s_rk4_mat_{frame_name}"] = solve_lmde(mm, t_span=[0, 0.5], y0=rho, method="RK4", max_dt=0.01).y[-1]
        mv.in_frame_basis = True
        oc = mv._operator_collection._operator_collection
        out[f"s_super_static_{frame_name}"] = oc.static_operator
        out[f"s_super_ops_{frame_name}"] = oc.operators
    # collection-level goldens (non-vectorised LindbladCollection with batch of rho)
    lc = LindbladCollection(static_hamiltonian=H0, hamiltonian_operators=Hs, static_dissipators=Lstat,
                            dissipator_operators=Ldyn)
    rho = np.array([Y[:, b].reshape(n, n, order="F") for b in range(B)])
    hc, dc = np.array([0.3, -0.7]), np.array([0.25, 0.4])
    out["s_hc"], out["s_dc"] = hc, dc
    out["s_coll_rhs"] = lc.evaluate_rhs(hc, dc, rho)
    out["s_coll_rhs_hamonly"] = LindbladCollection(static_hamiltonian=H0, hamiltonian_operators=Hs).evaluate_rhs(hc, None, rho)
    out["s_coll_rhs_statdis"] = LindbladCollection(static_hamiltonian=H0, static_dissipators=Lstat).evaluate_rhs(None, None, rho)
    vc = VectorizedLindbladCollection(static_hamiltonian=H0, hamiltonian_operators=Hs, static_dissipators=Lstat,
                                      dissipator_operators=Ldyn)
    out["s_vcoll_eval"] = vc.evaluate(hc, dc)
    out["s_vcoll_rhs"] = vc.evaluate_rhs(hc, dc, Y)

    # cfg3-like: n=27 (729), 3 ham ops, 6 static dissipators, B=4, frame = diag(H0) (1-d), expm, 3 steps
    n, K, B = 27, 3, 4
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 6, B, 2003)
    out["cfg3_check"] = checksum(H0, Hs, Ls, Y)
    mv = LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(sig),
                       static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=True)
    out["cfg3_expm_y"] = solve_lmde(mv, t_span=[0, 0.03], y0=Y, method="scipy_expm", max_dt=1e-2).y[-1]
    out["cfg3_rhs"] = mv(0.013, Y)
    save("lindblad", **out)


def gen_magnus():
    """scipy_expm_solver at Magnus orders 2 and 3 (solvers/fixed_step_solvers.py:80-108, 327-401): Hamiltonian
    model (no frame / full frame, matrix and vector y0, t_eval), a plain callable generator, and the small
    vectorised Lindblad system in its three frames; step sizes large enough that the commutator terms matter."""
    out = {}
    n, K, B = 5, 2, 3
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 11)
    out["h_check"] = checksum(H0, Hs, Y)
    for order in (2, 3):
        m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig))
        out[f"h_noframe_o{order}"] = solve_lmde(m, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.05,
                                                magnus_order=order).y[-1]
        m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig), rotating_frame=H0)
        out[f"h_full_o{order}"] = solve_lmde(m, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.05,
                                             magnus_order=order).y[-1]
        out[f"h_full_vec_o{order}"] = solve_lmde(m, t_span=[0, 0.5], y0=Y[:, 0], method="scipy_expm", max_dt=0.05,
                                                 magnus_order=order).y[-1]
        r = solve_lmde(m, t_span=[0.5, 0.0], y0=Y, method="scipy_expm", max_dt=0.04, magnus_order=order,
                       t_eval=[0.5, 0.31, 0.1])
        out[f"h_full_teval_back_o{order}"] = r.y
        # plain callable generator (non-commuting at different times)
        A = -1j * H0
        Bm = -1j * Hs[0]
        r = solve_lmde(lambda t: A * np.cos(t) + Bm * np.sin(2 * t), t_span=[0, 1.0], y0=np.eye(n, dtype=complex),
                       method="scipy_expm", max_dt=0.1, magnus_order=order)
        out[f"callable_o{order}"] = r.y[-1]
    # 17-dimensional system: wider than one DMMA tile, not a multiple of 8
    n, K, B = 17, 3, 6
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 41)
    out["h17_check"] = checksum(H0, Hs, Y)
    m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig), rotating_frame=H0)
    for order in (2, 3):
        out[f"h17_full_o{order}"] = solve_lmde(m, t_span=[0, 0.3], y0=Y, method="scipy_expm", max_dt=0.03,
                                               magnus_order=order).y[-1]
    # small vectorised Lindblad (the system of gen_lindblad)
    n, K, B = 3, 2, 4
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 4, B, 31)
    Lstat, Ldyn = Ls[:2], Ls[2:] + 0.02j * Ls[:2]
    dsig = [(0.3, 0.0, 0.0), (0.2, 0.11, 0.4)]
    out["l_check"] = checksum(H0, Hs, Ls, Y)
    for frame_name, frame in (("none", None), ("full", H0), ("diag", np.diag(H0).real)):
        mv = LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(sig),
                           static_dissipators=Lstat, dissipator_operators=Ldyn, dissipator_signals=sigs(dsig),
                           rotating_frame=frame, vectorized=True)
        for order in (2, 3):
            out[f"l_{frame_name}_o{order}"] = solve_lmde(mv, t_span=[0, 0.5], y0=Y, method="scipy_expm", max_dt=0.05,
                                                         magnus_order=order).y[-1]
    save("magnus", **out)


# ---------------------------------------------------------------------------------------------
MEASUREMENT_CASES = [
    # (name, subsystem_dims, measurement_subsystems, memory_slot_indices, num_memory_slots, max_outcome_level)
    ("two_transmons_both", [3, 3], [0, 1], [0, 1], None, 1),
    ("two_transmons_q1_slot2", [3, 3], [1], [2], 3, None),
    ("two_transmons_q0_levels", [3, 3], [0], [0], None, 2),
    ("mixed_dims_swapped_slots", [2, 3, 2], [0, 2], [1, 0], None, 1),
]


def measurement_system(dims, seed):
    """Duffing-like static Hamiltonian with weak exchange coupling (nearly diagonal, so that the dressed-state
    sorting of backend_utils.py:31-80 is well defined), one drive operator, random final states."""
    rng = np.random.default_rng(seed)
    n = int(np.prod(dims))
    levels = np.zeros((len(dims), n))
    for i in range(n):
        rem = i
        for s, d in enumerate(dims):
            levels[s, i] = rem % d
            rem //= d
    H0 = np.zeros((n, n), dtype=complex)
    for s in range(len(dims)):
        H0 += np.diag(2 * np.pi * ((4.8 + 0.31 * s) * levels[s] - 0.15 * levels[s] * (levels[s] - 1)))
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H0 = H0 + 0.02 * (A + A.conj().T)
    Hd = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Hd = (Hd + Hd.conj().T) / 2
    B = 5
    Y = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    Y[:, 0] = 0.0
    Y[0, 0] = 1.0  # one column: a basis state (many zero-probability outcomes)
    return H0, Hd, Y


def gen_measurement():
    """Row f4.  Reference code executed: RotatingFrame.state_out_of_frame, _get_lab_frame_static_hamiltonian,
    _get_dressed_state_decomposition, _get_memory_slot_probabilities (backend/backend_utils.py) in the order of
    _get_experiment_result (backend/dynamics_backend.py:846-866).  Statevector.probabilities_dict belongs to the
    absent qiskit package and is restated in oracle.numpy_oracle.subsystem_probabilities_dict."""
    from qiskit_dynamics.backend import backend_utils as bu
    out = {}
    tf = 1.7
    for ci, (name, dims, meas, slots, nslots, max_level) in enumerate(MEASUREMENT_CASES):
        H0, Hd, Y = measurement_system(dims, 500 + ci)
        model = HamiltonianModel(static_operator=H0, operators=[Hd], signals=[Signal(0.1, 4.8)], rotating_frame=H0)
        lab_h = bu._get_lab_frame_static_hamiltonian(model)
        evals, dressed = bu._get_dressed_state_decomposition(lab_h)
        dicts = []
        for b in range(Y.shape[1]):
            yf = np.array(model.rotating_frame.state_out_of_frame(t=tf, y=Y[:, b]))
            yf = dressed.conj().T @ yf
            yf = yf / np.linalg.norm(yf)
            pd = orc.subsystem_probabilities_dict(np.abs(yf) ** 2, dims, meas)
            dicts.append(bu._get_memory_slot_probabilities(pd, slots, num_memory_slots=nslots, max_outcome_value=max_level))
        labels = sorted(set().union(*[d.keys() for d in dicts]))
        P = np.array([[d.get(lab, 0.0) for d in dicts] for lab in labels])
        out[f"{name}_H0"], out[f"{name}_Hd"], out[f"{name}_Y"] = H0, Hd, Y
        out[f"{name}_lab_h"], out[f"{name}_dressed_evals"], out[f"{name}_dressed_states"] = lab_h, evals, dressed
        out[f"{name}_labels"] = np.array(labels)
        out[f"{name}_probs"] = P
    out["tf"] = np.array(tf)
    save("measurement", **out)

def gen_fullsize():
    """Column subsets of the FULL-SIZE BASELINE configurations solved by the unmodified reference: the GPU tests run the
    whole batch at the stated sizes (cfg2 n=32 B=1024 1000 steps sweep; cfg3 729 x 4096 expm T=0.2; cfg4 n=128 B=4096
    1000 steps; cfg5-like n=81, 8 DiscreteSignal channels, 8192 sweep points) and compare these columns.  Every column
    evolves independently in the reference (no inter-column term), so solving the chosen columns alone is
    what `results[:, cols]` of the full batch would hold."""
    import bench_workloads as W
    out = {}
    # cfg4: 32 columns of the 4096, 1000 RK4 steps
    H0, Hs, Y, sig = W.cfg4()
    cols = W.parity_columns(Y.shape[1])
    m = HamiltonianModel(static_operator=H0, operators=Hs, signals=sigs(sig), rotating_frame=H0)
    out["cfg4_cols"] = cols
    out["cfg4_check"] = checksum(H0, Hs, Y)
    out["cfg4_y"] = solve_lmde(m, t_span=[0, 1.0], y0=Y[:, cols], method="RK4", max_dt=W.MAX_DT).y[-1]
    # a ragged batch (4090 columns: the last octet of the tiling is partial) shares the operators; its tail columns
    colsr = np.arange(4080, 4090)
    out["cfg4_ragged_cols"] = colsr
    out["cfg4_ragged_y"] = solve_lmde(m, t_span=[0, 0.1], y0=Y[:, colsr], method="RK4", max_dt=W.MAX_DT).y[-1]
    # cfg2: 32 of the 1024 sweep points through Solver.solve's sequential list loop, 1000 RK4 steps
    H0, Hs, y0, per_col = W.cfg2()
    cols = W.parity_columns(len(per_col))
    s = Solver(static_hamiltonian=H0, hamiltonian_operators=Hs, rotating_frame=H0)
    res = s.solve(t_span=[0, 1.0], y0=y0, signals=[sigs(per_col[b]) for b in cols], method="RK4", max_dt=W.MAX_DT)
    out["cfg2_cols"] = cols
    out["cfg2_check"] = checksum(H0, Hs, y0)
    out["cfg2_y"] = np.stack([r.y[-1] for r in res], axis=-1)
    # cfg3: 16 of the 4096 density matrices, 20 exponential steps
    H0, Hs, Ls, Y, sig = W.cfg3()
    cols = W.parity_columns(Y.shape[1], count=16)[:16]
    mv = LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(sig),
                       static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=True)
    out["cfg3_cols"] = cols
    out["cfg3_check"] = checksum(H0, Hs, Ls, Y)
    out["cfg3_y"] = solve_lmde(mv, t_span=[0, 0.2], y0=Y[:, cols], method="scipy_expm", max_dt=1e-2).y[-1]
    # row f2 at the cfg3 system: NON-vectorised LindbladModel, RK4 (max_dt = 1e-3, 20 steps) on 8 of the 4096 density
    # matrices as a (l, n, n) batch -- batch = leading axis (models/operator_collections.py:506-510)
    colsm = cols[:8]
    n3 = H0.shape[0]
    rho = np.array([Y[:, b].reshape(n3, n3, order="F") for b in colsm])
    mm = LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=sigs(sig),
                       static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=False)
    out["cfg3_mat_cols"] = colsm
    out["cfg3_mat_rk4_y"] = solve_lmde(mm, t_span=[0, 0.02], y0=rho, method="RK4", max_dt=1e-3).y[-1]
    out["cfg3_mat_rhs"] = mm(0.013, rho)
    # time-dependent dissipators + full (non-diagonal) frame at dim 20, batch of 5
    H20, Hs20, Ls20, Y20, sig20 = orc.synthetic_lindblad(20, 2, 5, 5, 2020)
    Lst20, Ldy20 = Ls20[:2], Ls20[2:] * (1 + 0.3j)
    dsig20 = [(0.7, 0.0, 0.0), (0.4, 0.31, 0.2), (0.9, 0.05, -0.4)]
    rho20 = np.array([Y20[:, b].reshape(20, 20, order="F") for b in range(5)])
    m20 = LindbladModel(static_hamiltonian=H20, hamiltonian_operators=Hs20, hamiltonian_signals=sigs(sig20),
                        static_dissipators=Lst20, dissipator_operators=Ldy20, dissipator_signals=sigs(dsig20),
                        rotating_frame=H20, vectorized=False)
    out["l20_check"] = checksum(H20, Hs20, Ls20, Y20)
    out["l20_rk4_y"] = solve_lmde(m20, t_span=[0, 0.1], y0=rho20, method="RK4", max_dt=2e-3).y[-1]
    out["l20_rk4_single_y"] = solve_lmde(m20, t_span=[0, 0.1], y0=rho20[1], method="RK4", max_dt=2e-3).y[-1]
    out["l20_rhs"] = m20(0.37, rho20)
    # cfg5-like: 32 of 8192 sweep points, DiscreteSignal Gaussian-square tables, max_dt = sample width (every stage
    # time on a bin edge), 64 RK4 steps; final states and the memory-slot probabilities of the reference's
    # post-processing chain (backend_utils)
    from qiskit_dynamics.backend import backend_utils as bu
    nsim, nsamp = 8192, 64
    H0, ops, freqs = W.cfg5_system()
    cols = W.parity_columns(nsim)
    s = Solver(static_hamiltonian=H0, hamiltonian_operators=list(ops), rotating_frame=H0)
    y0 = np.zeros(H0.shape[0], dtype=complex)
    y0[0] = 1.0
    lists = [[DiscreteSignal(dt=W.CFG5_DT, samples=smp, carrier_freq=float(freqs[j]), phase=ph)
              for j, (smp, ph) in enumerate(W.cfg5_point(int(k), nsim, nsamp))] for k in cols]
    tf = nsamp * W.CFG5_DT
    res = s.solve(t_span=[0, tf], y0=y0, signals=lists, method="RK4", max_dt=W.CFG5_DT)
    finals = np.stack([r.y[-1] for r in res], axis=-1)
    dims, msub, mslots = W.cfg5_measurement()
    lab_h = bu._get_lab_frame_static_hamiltonian(s.model)
    _, dressed = bu._get_dressed_state_decomposition(lab_h)
    dicts = []
    for b in range(finals.shape[1]):
        yf = np.array(s.model.rotating_frame.state_out_of_frame(t=tf, y=finals[:, b]))
        yf = dressed.conj().T @ yf
        yf = yf / np.linalg.norm(yf)
        pd = orc.subsystem_probabilities_dict(np.abs(yf) ** 2, dims, msub)
        dicts.append(bu._get_memory_slot_probabilities(pd, mslots, max_outcome_value=1))
    labels = sorted(set().union(*[d.keys() for d in dicts]))
    out["cfg5_cols"] = cols
    out["cfg5_check"] = checksum(H0, ops)
    out["cfg5_y"] = finals
    out["cfg5_labels"] = np.array(labels)
    out["cfg5_probs"] = np.array([[d.get(lab, 0.0) for d in dicts] for lab in labels])
    # the same system swept over 65 536 points (BASELINE configs[4] at size; bench.py's cfg5 record): 8 points
    nbig = 65536
    pts = np.array([0, 1, 4097, 8191, nbig // 3, nbig // 2, nbig - 9, nbig - 1])
    lists = [[DiscreteSignal(dt=W.CFG5_DT, samples=smp, carrier_freq=float(freqs[j]), phase=ph)
              for j, (smp, ph) in enumerate(W.cfg5_point(int(k), nbig, nsamp))] for k in pts]
    res = s.solve(t_span=[0, tf], y0=y0, signals=lists, method="RK4", max_dt=W.CFG5_DT)
    dicts = []
    for r in res:
        yf = np.array(s.model.rotating_frame.state_out_of_frame(t=tf, y=r.y[-1]))
        yf = dressed.conj().T @ yf
        yf = yf / np.linalg.norm(yf)
        dicts.append(bu._get_memory_slot_probabilities(orc.subsystem_probabilities_dict(np.abs(yf) ** 2, dims, msub), mslots,
                                                       max_outcome_value=1))
    labels_big = sorted(set().union(*[d.keys() for d in dicts]))
    out["cfg5_big_points"] = pts
    out["cfg5_big_labels"] = np.array(labels_big)
    out["cfg5_big_probs"] = np.array([[d.get(lab, 0.0) for d in dicts] for lab in labels_big])
    save("fullsize", **out)


if __name__ == "__main__":
    only = set(sys.argv[1:])  # e.g. `make_golden.py magnus` regenerates one file
    if only:
        for name in only:
            globals()["gen_" + name]()
        raise SystemExit(0)
    gen_magnus()
    gen_measurement()
    gen_collection()
    gen_frame()
    gen_signals()
    gen_step_grid()
    gen_hamiltonian_model()
    gen_rk4_solves()
    gen_lindblad()
    gen_fullsize()
