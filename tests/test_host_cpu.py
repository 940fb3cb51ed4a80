"""CPU-only tests: host logic of the drop-in surface, the C-ABI library's symbols, and the
world_size-2 sharding path on gloo.  No compute call is made (there is no GPU here and the product
has no CPU fallback -- which is itself asserted)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import numpy_oracle as orc


@pytest.fixture(scope="module")
def qd():
    import __graft_entry__ as ge
    ge.build()
    import qiskit_dynamics_b200 as q
    return q


def test_library_exports_every_declared_symbol(qd):
    header = open(os.path.join(ROOT, "include", "qdb.h")).read()
    declared = set(re.findall(r"\b(qdb_[a-z0-9_]+)\s*\(", header))
    declared -= {"qdb_c128"}
    assert len(declared) >= 12
    lib = ctypes.CDLL(qd._abi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/qdb.h but not exported"
    assert declared == set(qd._abi.SIGNATURES), (declared ^ set(qd._abi.SIGNATURES))
    l = qd._abi.lib()
    assert l.qdb_version() >= 100
    assert l.qdb_npad(27) == 32 and l.qdb_npad(128) == 128 and l.qdb_packed_elems(5) == 8 * 16
    assert qd._abi.workspace_bytes(qd._abi.WS_RK4, 128, 8, 4096, 10) >= 21 * 128 * 128 * 16
    assert l.qdb_last_error_string() is not None


def test_abi_argument_validation_without_gpu(qd):
    """Invalid arguments are rejected before any CUDA call (negative return, message set)."""
    l = qd._abi.lib()
    rc = l.qdb_pack_operators(0, 1, None, None, None)
    assert rc == -1 and b"qdb_pack_operators" in l.qdb_last_error_string()
    rc = l.qdb_generator_c128(4, 0, 1, 0, None, None, None, 0, None, None, 1.0, None, None)
    assert rc == -1
    rc = l.qdb_rk4_steps_c128(4, 1, 2, 3, None, None, None, None, None, 7, 0, None, None, 0.1, None, 2, None, 0, None)
    assert rc == -1 and b"sig_mode" in l.qdb_last_error_string()
    assert l.qdb_rk4_steps_c128(4, 1, 0, 3, None, None, None, None, None, 0, 0, None, None, 0.1, None, 0, None, 0, None) == 0  # empty batch
    assert l.qdb_expm_steps_c128(4, 0, 1, 0, None, None, None, None, None, None, 0.1, None, 1, None, 0, None) == 0  # zero steps
    assert l.qdb_magnus_steps_c128(4, 0, 1, 0, 3, None, None, None, None, None, None, 0.1, None, 1, None, 0, None) == 0  # zero steps
    rc = l.qdb_magnus_steps_c128(4, 0, 1, 2, 4, None, None, None, None, None, None, 0.1, None, 1, None, 0, None)
    assert rc == -1 and b"Only magnus_order 1, 2, and 3" in l.qdb_last_error_string()
    assert l.qdb_magnus_terms_c128(4, 0, None, 0.1, 1.0, None, None, 0, None) == -1
    assert qd._abi.workspace_bytes(qd._abi.WS_MAGNUS, 8, 2, 4, 5) >= 17 * 64 * 16 + 8 * 4 * 16 + 15 * 8


def test_magnus_host_tables():
    """Node offsets, node-time tables and the norm bound that picks the squarings (host side of row a9)."""
    from qiskit_dynamics_b200.solvers.fixed_step import magnus_nodes, magnus_norm_bound
    from qiskit_dynamics_b200 import QiskitError
    for order in (1, 2, 3):
        assert np.array_equal(magnus_nodes(order), orc.magnus_node_offsets(order))  # same expressions as the reference
    with pytest.raises(QiskitError, match="Only magnus_order 1, 2, and 3"):
        magnus_nodes(0)
    rng = np.random.default_rng(5)
    for order in (1, 2, 3):
        for _ in range(10):
            gs = [rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6)) for _ in range(order)]
            it = iter(gs)
            omega = orc.magnus_propagator(lambda t: next(it), 0.0, -0.3, order, expm_func=lambda x: x)
            b = np.array([[np.linalg.norm(g_, 1) for g_ in gs]])
            assert np.linalg.norm(omega, 1) <= magnus_norm_bound(b, -0.3, order)[0] * (1 + 1e-12)


def test_no_cpu_fallback(qd):
    m = qd.HamiltonianModel(static_operator=np.diag([1.0, -1.0]), operators=[np.array([[0, 1], [1, 0]])],
                            signals=[qd.Signal(1.0, 1.0)], rotating_frame=np.diag([1.0, -1.0]))
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(qd._abi.QdbError, match="no CPU fallback"):
        m(0.1, np.array([1.0, 0.0]))
    with pytest.raises(qd._abi.QdbError):
        qd.solve_lmde(m, t_span=[0, 1], y0=np.array([1.0, 0.0]), method="RK4", max_dt=0.1)
    # nothing in the product imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qiskit_dynamics_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("numpy_oracle", "oracle") or "import" not in [ln for ln in src.splitlines() if "oracle" in ln][0], f


def test_signals_match_reference_golden(qd):
    g = load_golden("signals")
    ts, samples = g["ts"], g["samples"]
    d1 = qd.DiscreteSignal(dt=0.1, samples=samples, carrier_freq=1.3, phase=0.2)
    d2 = qd.DiscreteSignal(dt=0.1, samples=samples, start_time=1.0, carrier_freq=0.0)
    s1 = qd.Signal(0.7, 2.0, 0.4)
    s2 = qd.Signal(lambda t: np.exp(-t**2) * (1 + 0.5j), carrier_freq=0.9, phase=-1.1)
    s3 = qd.Signal(1.5)
    tol = dict(rtol=0, atol=1e-14)
    np.testing.assert_allclose(d1(ts), g["d1"], **tol)
    np.testing.assert_allclose(d1.complex_value(ts), g["d1_cv"], **tol)
    np.testing.assert_allclose(d2(ts), g["d2"], **tol)
    np.testing.assert_allclose(s1(ts), g["s1"], **tol)
    np.testing.assert_allclose(s2(ts), g["s2"], **tol)
    np.testing.assert_allclose(s2.complex_value(ts), g["s2_cv"], **tol)
    np.testing.assert_allclose(s3(ts), g["s3"], **tol)
    np.testing.assert_allclose((s1 + s2)(ts), g["ssum"], **tol)
    np.testing.assert_allclose((s1 * s2)(ts), g["sprod"], **tol)
    np.testing.assert_allclose((d1 * d1)(ts), g["dprod"], **tol)
    sl = qd.SignalList([s1, s2, s3, d1, d2, s1 + s2, 2.0])
    np.testing.assert_allclose(sl(ts), g["siglist"], **tol)
    np.testing.assert_allclose(sl(0.123), g["siglist_scalar"], **tol)
    np.testing.assert_allclose(sl.complex_value(ts), g["siglist_cv"], **tol)
    np.testing.assert_allclose(sl.drift, g["drift"], **tol)
    assert sl.table(ts).shape == (len(ts), 7) and sl.table(ts).flags["C_CONTIGUOUS"]
    np.testing.assert_allclose(d1.conjugate().complex_value(ts), g["conj_d1"], **tol)
    # bin-edge semantics on accumulated stage times: identical sample choice
    d3 = qd.DiscreteSignal(dt=1 / 4.5, samples=g["dsamp"], carrier_freq=0.4)
    assert np.array_equal(d3.envelope(g["tacc"]), g["d3_env_acc"])
    np.testing.assert_allclose(d3(g["tacc"]), g["d3_acc"], **tol)
    # algebra closure / types
    assert s3.is_constant and not s1.is_constant
    assert isinstance(d1 + d1, qd.DiscreteSignalSum) and isinstance(2.0 * d1, qd.DiscreteSignalSum)
    assert isinstance(s1 * s2, qd.SignalSum) and len(s1 * s2) == 2
    np.testing.assert_allclose((s1 - s2)(ts), g["s1"] - g["s2"], **tol)
    np.testing.assert_allclose((-s1)(ts), -g["s1"], **tol)
    np.testing.assert_allclose((3 + s1)(ts), 3 + g["s1"], **tol)
    np.testing.assert_allclose((s1 + s2).flatten()(ts), g["ssum"], rtol=0, atol=1e-13)
    ds = qd.DiscreteSignal.from_Signal(s2, dt=0.1, n_samples=10)
    assert ds.duration == 10 and ds.dt == 0.1
    np.testing.assert_allclose(ds.samples, s2.envelope(0.05 + 0.1 * np.arange(10)), **tol)
    dd = qd.DiscreteSignal(dt=0.5, samples=[1.0, 2.0])
    dd.add_samples(3, [5.0])
    np.testing.assert_allclose(dd.samples, [1, 2, 0, 5])
    with pytest.raises(qd.QiskitError):
        dd.add_samples(1, [1.0])
    assert len(sl[[0, 2]]) == 2 and isinstance(sl[1], qd.SignalSum)
    with pytest.raises(qd.QiskitError):
        qd.SignalSum("not a signal")


def test_step_grid_matches_reference_golden(qd):
    from qiskit_dynamics_b200.solvers import get_fixed_step_sizes, merge_t_args, stage_time_grid
    g = load_golden("step_grid")
    for i in range(int(g["ncases"])):
        ev = g[f"eval{i}"] if bool(g[f"has_eval{i}"]) else None
        t, h, n = get_fixed_step_sizes(g[f"span{i}"], ev, float(g[f"maxdt{i}"]))
        assert np.array_equal(t, g[f"t{i}"]) and np.array_equal(h, g[f"h{i}"]) and np.array_equal(n, g[f"n{i}"])
    for bad in ([0.5, 1.5], [0.7, 0.5], [[0.5]]):
        with pytest.raises(ValueError):
            merge_t_args([0, 1], bad)
    # stage grid == the reference loop's accumulated floats
    t0, h, S = 0.3, (1 / 4.5) / 2, 25
    grid = stage_time_grid(t0, h, S)
    t = t0
    for i in range(S):
        assert grid[2 * i] == t and grid[2 * i + 1] == t + 0.5 * h and grid[2 * i + 2] == t + h
        t = t + h


def test_model_construction_and_error_conventions(qd):
    g = load_golden("hamiltonian_model")
    H0, Hs = g["H0"], g["Hs"]
    m = qd.HamiltonianModel(static_operator=H0, operators=Hs, rotating_frame=np.diag(H0).real, in_frame_basis=True)
    np.testing.assert_allclose(m._operator_collection.operators.numpy(), g["ops_diag"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(m._operator_collection.static_operator.numpy(), g["stat_diag"], rtol=0, atol=1e-14)
    m0 = qd.HamiltonianModel(static_operator=H0, operators=Hs)
    np.testing.assert_allclose(m0._operator_collection.operators.numpy(), g["ops_none"], rtol=0, atol=1e-14)
    assert m.dim == 8 and m.in_frame_basis and m.signals is None
    m.signals = [1.0, 2.0, qd.Signal(1.0, 3.0)]
    assert isinstance(m.signals, qd.SignalList) and len(m.signals) == 3
    np.testing.assert_allclose(m.rotating_frame.frame_freqs.numpy(), np.diag(H0).real)
    for bad, exc in (
        (lambda: qd.HamiltonianModel(), qd.QiskitError),
        (lambda: qd.HamiltonianModel(static_operator=np.array([[0, 1], [0, 0]])), qd.QiskitError),
        (lambda: qd.HamiltonianModel(operators=Hs, signals=[1.0]), qd.QiskitError),
        (lambda: qd.HamiltonianModel(static_operator=H0, signals=[1.0]), qd.QiskitError),
        (lambda: qd.HamiltonianModel(operators=Hs, signals="x"), qd.QiskitError),
        (lambda: qd.RotatingFrame(np.array([[1.0, 2.0], [3.0, 4.0]])), qd.QiskitError),
        (lambda: qd.LindbladModel(), qd.QiskitError),
        (lambda: qd.LindbladModel(static_hamiltonian=np.array([[0, 1], [0, 0]])), qd.QiskitError),
        (lambda: qd.Solver(hamiltonian_operators=Hs, dt=0.1), qd.QiskitError),
        (lambda: qd.OperatorCollection(operators=Hs, array_library="jax"), qd.QiskitError),
    ):
        with pytest.raises(exc):
            bad()
    # vectorised Lindblad superoperators are built at construction (no frame -> pure torch set-up)
    gl = load_golden("lindblad")
    lm = qd.LindbladModel(static_hamiltonian=gl["s_H0"], hamiltonian_operators=gl["s_Hs"], static_dissipators=gl["s_Lstat"],
                          dissipator_operators=gl["s_Ldyn"], vectorized=True)
    oc = lm._operator_collection._operator_collection
    np.testing.assert_allclose(oc.static_operator.numpy(), gl["s_super_static_none"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(oc.operators.numpy(), gl["s_super_ops_none"], rtol=0, atol=1e-13)
    lm1 = qd.LindbladModel(static_hamiltonian=gl["s_H0"], hamiltonian_operators=gl["s_Hs"], static_dissipators=gl["s_Lstat"],
                           dissipator_operators=gl["s_Ldyn"], vectorized=True, rotating_frame=np.diag(gl["s_H0"]).real)
    np.testing.assert_allclose(lm1._operator_collection._operator_collection.static_operator.numpy(),
                               gl["s_super_static_diag"], rtol=0, atol=1e-13)
    lam = np.diag(gl["s_H0"]).real
    np.testing.assert_allclose(lm1._frame_freqs().numpy(), (lam[:, None] - lam[None, :]).flatten(order="F"))
    with pytest.raises(qd.QiskitError):
        lm.signals = ([1.0], None)
    # Solver argument plumbing
    from qiskit_dynamics_b200.solvers.solver_classes import setup_args_lists, t_span_to_list, _y0_to_list, _signals_to_list
    lists, multi = setup_args_lists([[0, 1], [np.ones(2), np.zeros(2)], None], ["t_span", "y0", "signals"],
                                    [t_span_to_list, _y0_to_list, _signals_to_list])
    assert multi and [len(x) for x in lists] == [2, 2, 2]
    with pytest.raises(qd.QiskitError):
        setup_args_lists([[[0, 1], [0, 2], [0, 3]], [np.ones(2), np.zeros(2)], None], ["t_span", "y0", "signals"],
                         [t_span_to_list, _y0_to_list, _signals_to_list])


def _npy_floor_divide(a, b):
    """Python restatement of npy_floor_divide in csrc/signals.cu (NumPy's npy_divmod), same operation order."""
    import math
    if b == 0.0:
        return math.copysign(math.inf, a) if a != 0 else math.nan
    mod = math.fmod(a, b)
    div = (a - mod) / b
    if mod != 0.0 and ((b < 0.0) != (mod < 0.0)):
        div -= 1.0
    if div != 0.0:
        fd = math.floor(div)
        if div - fd > 0.5:
            fd += 1.0
        return fd
    return math.copysign(0.0, a / b)


def test_bin_index_rule_is_numpy_floor_division():
    """The device bin selector restates NumPy's float floor division, which the reference uses
    (signals/signals.py:304-308) and which differs from floor(a / b) on bin edges (1.0 // 0.1 == 9)."""
    rng = np.random.default_rng(0)
    assert np.float64(1.0) // np.float64(0.1) == 9.0 and np.floor(1.0 / 0.1) == 10.0
    cases = [(1.0, 0.1), (0.3, 0.1), (0.7, 0.1), (-0.3, 0.1), (2.0, 0.25), (0.0, 0.1), (-0.0, 0.1), (1e-300, 0.1),
             (5.0, -0.5), (-5.0, -0.5), (0.30000000000000004, 0.1)]
    dts = [0.1, 0.2222222222222222, 1 / 4.5, 1e-3, 0.25, 3.0]
    for dt in dts:  # accumulated stage times t <- t + h landing on bin edges
        t = 0.0
        for _ in range(200):
            cases.append((t, dt))
            cases.append((t + dt / 2, dt))
            t = t + dt / 2
    cases += [(float(a), float(b)) for a, b in zip(rng.uniform(-50, 50, 2000), rng.uniform(0.01, 3, 2000))]
    for a, b in cases:
        want = np.float64(a) // np.float64(b)
        got = _npy_floor_divide(a, b)
        assert got == want and np.signbit(got) == np.signbit(want), (a, b, got, want)


def test_signal_program_compile(qd):
    """Flattening of SignalLists into the device term arrays (row f3)."""
    from qiskit_dynamics_b200.signals import compile_signal_program
    sl = qd.SignalList([qd.Signal(0.3, 5.0, 0.1), qd.DiscreteSignal(0.1, np.arange(5) + 1j, carrier_freq=2.0) + 0.5, 1.0])
    p = compile_signal_program(sl)
    assert p.columns == 0 and p.num_channels == 3
    assert list(p.chan) == [0, 1, 1, 2] and list(p.samp_len) == [-1, 5, -1, -1] and list(p.samp_off) == [0, 1, 6, 7]
    np.testing.assert_array_equal(p.freq, [5.0, 2.0, 0.0, 0.0])
    np.testing.assert_array_equal(p.samples, np.concatenate([[0.3], np.arange(5) + 1j, [0.5], [1.0]]))
    # sweep: amplitude differs -> per-column samples, shared parameters
    lists = [qd.SignalList([qd.Signal(0.3 * (b + 1), 5.0, 0.1), qd.DiscreteSignal(0.1, (np.arange(5) + 1j) * b, carrier_freq=2.0)])
             for b in range(3)]
    p = compile_signal_program(lists)
    assert p.columns == 3 and not p.params_per_column and p.samples.shape == (3, 6)
    # frequency sweep -> per-column parameters, shared samples
    lists = [qd.SignalList([qd.Signal(0.3, 5.0 + b, 0.1)]) for b in range(4)]
    p = compile_signal_program(lists)
    assert p.params_per_column and p.freq.shape == (1, 4) and p.samples.shape == (1,)
    # arbitrary Python envelopes and structure mismatches stay on the host path
    assert compile_signal_program(qd.SignalList([qd.Signal(lambda t: t)])) is None
    assert compile_signal_program([qd.SignalList([qd.Signal(1.0)]), qd.SignalList([qd.DiscreteSignal(0.1, [1.0])])]) is None


def test_memory_slot_outcome_map_matches_oracle(qd):
    """The product's basis-state -> outcome table (host set-up of row f4) against the oracle's dictionary
    pipeline (pinned by the measurement fixtures) on every basis state."""
    from conftest import MEASUREMENT_CASES
    from oracle import numpy_oracle as orc
    from qiskit_dynamics_b200.measurement import memory_slot_outcome_map
    for name, dims, meas, slots, nslots, max_level in MEASUREMENT_CASES + [("wide", [2, 2, 3], [2, 0, 1], [0, 3, 1], 5, None)]:
        labels, outcome_of = memory_slot_outcome_map(dims, meas, slots, nslots, max_level)
        n = int(np.prod(dims))
        assert outcome_of.shape == (n,) and outcome_of.dtype == np.int32 and labels == sorted(labels)
        for i in range(n):
            probs = np.zeros(n)
            probs[i] = 1.0
            d = orc.memory_slot_probabilities(orc.subsystem_probabilities_dict(probs, dims, meas), slots, nslots, max_level)
            assert d == {labels[outcome_of[i]]: 1.0}, (name, i, d)


def test_shard_bounds():
    from qiskit_dynamics_b200 import distributed as D
    for B in (1, 7, 512, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(B, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world_size_2_gather():
    """Batch sharding + the single final all-gather, two processes on gloo (CPU tensors)."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", script],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GLOO_OK rank=0" in res.stdout and "GLOO_OK rank=1" in res.stdout


def test_bench_workload_matches_the_test_generator():
    """bench.py builds its inputs without importing oracle/; they must be the numbers the parity tests use."""
    sys.path.insert(0, ROOT)
    import bench
    for a, b in zip(bench.workload(16, 3, 5, 2004), orc.synthetic_schrodinger(16, 3, 5, 2004)):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    # the only code of bench.py that imports anything under oracle/ is reference_solver (the CPU legs): the B200 arm
    # checks its results against committed reference fixtures, never against the oracle
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    offenders = []
    for fn in [x for x in ast.walk(tree) if isinstance(x, ast.FunctionDef)]:
        for node in ast.walk(fn):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            if any(nm.split(".")[0] == "oracle" for nm in names) and fn.name != "reference_solver":
                offenders.append(fn.name)
    assert not offenders, f"oracle imported outside bench.reference_solver: {offenders}"
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in ast.dump(n) for n in top)


def test_signal_program_from_plain_lists(qd):
    """Large sweeps hand Solver.solve plain lists of elementary signals; compiling them directly must give the same
    device program as the SignalList route (which wraps every channel of every simulation into a SignalSum)."""
    from qiskit_dynamics_b200.signals import SignalList, compile_signal_program
    rng = np.random.default_rng(3)
    lists = [[qd.DiscreteSignal(dt=0.2, samples=rng.standard_normal(6) + 1j * rng.standard_normal(6), start_time=0.1 * (j % 2),
                                carrier_freq=1.0 + j, phase=0.3 * b) for j in range(3)] for b in range(7)]
    fast, slow = compile_signal_program(lists), compile_signal_program([SignalList(l) for l in lists])
    for name in ("chan", "samp_len", "dt", "t0", "freq", "phase"):
        assert np.array_equal(getattr(fast, name), getattr(slow, name)), name
    # the sample stores may be laid out differently (the plain-list route keeps the trailing zero of every signal): what
    # the kernel indexes -- samp_len samples from samp_off on, per channel and column -- must be the same
    for j in range(3):
        blk = lambda p: np.atleast_2d(p.samples)[:, p.samp_off[j]:p.samp_off[j] + p.samp_len[j]]  # noqa: E731
        assert np.array_equal(blk(fast), blk(slow)), j
    assert (fast.num_channels, fast.columns) == (slow.num_channels, slow.columns) == (3, 7)
    const = [[qd.Signal(0.3 * (b + 1), 1.0 + j, 0.1) for j in range(2)] for b in range(4)]
    fast, slow = compile_signal_program(const), compile_signal_program([SignalList(l) for l in const])
    for name in ("chan", "samp_len", "samp_off", "dt", "t0", "freq", "phase", "samples"):
        assert np.array_equal(getattr(fast, name), getattr(slow, name)), name
    assert compile_signal_program([[qd.Signal(lambda t: t, 1.0)]]) is None      # Python envelope: host path
    assert compile_signal_program([[qd.Signal(1.0)], [qd.DiscreteSignal(0.1, [1.0, 2.0])]]) is None  # structure differs


def test_integration_doc_matches_the_abi(qd):
    """The ctypes stub shown in INTEGRATION.md lists the same number of arguments as the binding the package uses."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    found = dict(re.findall(r"_lib\.(qdb_[a-z0-9_]+)\.argtypes = \[([^\]]*)\]", text))
    assert len(found) >= 10
    for name, args in found.items():
        assert name in qd._abi.SIGNATURES, name
        assert len([a for a in args.split(",") if a.strip()]) == len(qd._abi.SIGNATURES[name][1]), name


def test_signal_program_pads_ragged_pulse_sweeps(qd):
    """Simulations whose pulses have different sample counts compile to the program of the explicitly zero-padded
    signals (a DiscreteSignal is zero past its end), and evaluate like them on the host."""
    from qiskit_dynamics_b200.signals import SignalList, compile_signal_program
    rng = np.random.default_rng(8)
    lens = [[5, 3], [2, 7], [5, 7], [1, 1]]
    raw = [[rng.standard_normal(n_) + 1j * rng.standard_normal(n_) for n_ in row] for row in lens]
    mk = lambda samples, j, b: qd.DiscreteSignal(dt=0.2, samples=samples, start_time=0.1 * j, carrier_freq=1.0 + j, phase=0.2 * b)  # noqa: E731
    ragged = [[mk(raw[b][j], j, b) for j in range(2)] for b in range(4)]
    target = [5, 7]
    padded = [[mk(np.concatenate([raw[b][j], np.zeros(target[j] - len(raw[b][j]))]), j, b) for j in range(2)] for b in range(4)]
    a, c = compile_signal_program(ragged), compile_signal_program(padded)
    assert a is not None and c is not None
    for name in ("chan", "samp_len", "samp_off", "dt", "t0", "freq", "phase", "samples"):
        assert np.array_equal(getattr(a, name), getattr(c, name)), name
    assert list(a.samp_len) == target
    ts = np.linspace(-0.3, 2.5, 57)
    for b in range(4):
        assert np.array_equal(SignalList(ragged[b])(ts), SignalList(padded[b])(ts))
