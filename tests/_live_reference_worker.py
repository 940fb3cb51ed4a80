"""Worker for test_live_reference_cpu.py: imports the UNMODIFIED reference (through oracle/ref_shim.py) next to this
package and compares the host-side pieces of the path -- signals (row a6), step grids (row a8), Magnus nodes (row a9) --
on randomised inputs.  Runs in its own process because the shim installs import hooks for the absent qiskit package."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.ref_shim as shim  # noqa: E402

assert shim.reference_available()
import qiskit_dynamics as ref  # noqa: E402
from qiskit_dynamics.signals import DiscreteSignalSum as RDSS, SignalList as RSL, SignalSum as RSS  # noqa: E402
from qiskit_dynamics.solvers.fixed_step_solvers import get_fixed_step_sizes as r_steps  # noqa: E402
from qiskit_dynamics.solvers.solver_utils import merge_t_args as r_merge  # noqa: E402

import qiskit_dynamics_b200 as our  # noqa: E402
from qiskit_dynamics_b200.signals import DiscreteSignalSum as ODSS, SignalList as OSL, SignalSum as OSS  # noqa: E402
from qiskit_dynamics_b200.solvers.fixed_step import get_fixed_step_sizes as o_steps, merge_t_args as o_merge  # noqa: E402

rng = np.random.default_rng(2024)
checks = 0


def same(a, b, tol=1e-13):
    global checks
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.allclose(a, b, rtol=0, atol=tol), float(np.max(np.abs(a - b)))
    checks += 1


for trial in range(40):
    ts = np.concatenate([rng.uniform(-0.5, 3.0, 12), [0.0, 0.1, 0.2, 0.30000000000000004, 1.0]])
    nu, ph, amp = float(rng.uniform(0, 3)), float(rng.uniform(-2, 2)), complex(rng.standard_normal(), rng.standard_normal())
    dt = float(rng.choice([0.1, 0.05, 1 / 4.5, 0.37]))
    nsamp = int(rng.integers(1, 9))
    samples = rng.standard_normal(nsamp) + 1j * rng.standard_normal(nsamp)
    t0 = float(rng.choice([0.0, 0.1, -0.3, 0.7]))
    env = lambda t, a=amp: a * np.exp(-np.asarray(t) ** 2)  # noqa: E731
    pairs = []
    for mod, SL in ((ref, RSL), (our, OSL)):
        s_const = mod.Signal(amp, nu, ph)
        s_fun = mod.Signal(env, carrier_freq=0.5 * nu, phase=-ph)
        d = mod.DiscreteSignal(dt=dt, samples=samples, start_time=t0, carrier_freq=nu, phase=ph)
        d2 = mod.DiscreteSignal(dt=dt, samples=samples[::-1].copy(), start_time=t0, carrier_freq=0.3, phase=0.0)
        objs = {
            "const": s_const, "fun": s_fun, "disc": d, "sum": s_const + s_fun, "prod": s_const * s_fun, "dsum": d + d2, "dprod": d * d2,
            "scaled": 2.5 * d, "neg": -s_fun, "diff": d - s_const, "conj": d.conjugate(),
            "sampled": mod.DiscreteSignal.from_Signal(s_fun, dt=0.2, n_samples=7, start_time=0.1),
            "sampled_carrier": mod.DiscreteSignal.from_Signal(s_const, dt=0.2, n_samples=5, sample_carrier=True),
        }
        lst = SL([s_const, s_fun, d, s_const + d, 1.5])
        pairs.append((objs, lst))
    (ro, rl), (oo, ol) = pairs
    for key in ro:
        same(ro[key](ts), oo[key](ts))
        same(ro[key].complex_value(ts), oo[key].complex_value(ts))
        same(ro[key](0.123), oo[key](0.123))
    same(rl(ts), ol(ts))
    same(rl.complex_value(ts), ol.complex_value(ts))
    same(rl.drift, ol.drift)
    same(rl.flatten()(ts), ol.flatten()(ts))
    same(len(rl), len(ol))
    rs, os_ = RSS(ro["const"], ro["fun"]), OSS(oo["const"], oo["fun"])
    same(RDSS.from_SignalSum(rs, dt=0.15, n_samples=6, start_time=0.05)(ts), ODSS.from_SignalSum(os_, dt=0.15, n_samples=6, start_time=0.05)(ts))
    same(rs.flatten()(ts), os_.flatten()(ts))
    # step grids: random spans, t_eval inside, both directions, step sizes that divide the intervals exactly or not
    a, b = sorted(rng.uniform(-1, 2, 2))
    span = [a, b] if trial % 2 else [b, a]
    max_dt = float(rng.choice([(b - a) / 7, (b - a) / 7 * (1 + 1e-15), 0.013, 10.0]))
    te = None
    if trial % 3:
        pts = np.sort(rng.uniform(a, b, int(rng.integers(1, 5))))
        te = pts if span[0] < span[1] else pts[::-1]
    for x, y in zip(r_steps(np.array(span), te, max_dt), o_steps(span, te, max_dt)):
        assert np.array_equal(np.asarray(x), np.asarray(y)), (span, te, max_dt)
        checks += 1
    assert np.array_equal(np.asarray(r_merge(np.array(span), te)), np.asarray(o_merge(np.array(span), te)))

# error conventions of merge_t_args
for bad in ([[0, 1], [[0.1, 0.2]]], [[0, 1], [1.5]], [[0, 1], [0.5, 0.2]]):
    for fn in (r_merge, o_merge):
        try:
            fn(np.array(bad[0]), np.array(bad[1]))
        except ValueError:
            checks += 1
        else:
            raise AssertionError(("no ValueError", fn, bad))
print(f"LIVE_REFERENCE_OK checks={checks}")
