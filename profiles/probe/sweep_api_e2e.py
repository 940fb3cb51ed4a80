"""cfg5-style sweep end to end through the PUBLIC API: Solver.solve(t_span, y0, signals=[one list of DiscreteSignals per
simulation], method="RK4", max_dt=dt) -- host-side compilation of the signal lists, device signal table, one sweep-mode
launch per chunk, results split per simulation -- against the kernel-only time of the same sweep.  One JSON line.

    python profiles/probe/sweep_api_e2e.py [nsim] [n_levels] [n_samples]
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qiskit_dynamics_b200 as qd
from qiskit_dynamics_b200 import _abi as abi

nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nsamp = int(sys.argv[3]) if len(sys.argv) > 3 else 64
nq = 4
n = dim ** nq
a = np.diag(np.sqrt(np.arange(1, dim)), 1).astype(complex); N = a.conj().T @ a; I = np.eye(dim, dtype=complex)
def op(single, k):
    mats = [I] * nq; mats[nq - 1 - k] = single
    out = mats[0]
    for m in mats[1:]: out = np.kron(out, m)
    return out
w = 2 * np.pi * np.array([5.0, 5.1, 4.9, 5.05]); alpha, J = 2 * np.pi * -0.33, 2 * np.pi * 0.002
H0 = sum(w[k] * op(N, k) + 0.5 * alpha * op(N @ (N - I), k) for k in range(nq))
H0 = H0 + sum(J * (op(a, k) @ op(a.conj().T, k + 1) + op(a.conj().T, k) @ op(a, k + 1)) for k in range(nq - 1))
drives = [2 * np.pi * 0.02 * op(a + a.conj().T, k) for k in range(nq)]
ops = drives + drives  # 4 drive + 4 control channels
freqs = list(w / (2 * np.pi)) + list(np.roll(w, 1) / (2 * np.pi))
solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=ops, rotating_frame=H0)
dt = 0.222
t = (np.arange(nsamp) + 0.5) * dt
def envelope(amp, width):
    c, rise = t[-1] / 2 + dt / 2, 0.15 * nsamp * dt
    return amp * np.exp(-0.5 * (np.clip((np.abs(t - c) - width / 2) / rise, 0.0, None)) ** 2).astype(complex)
t0 = time.perf_counter()
signals = []
for k in range(nsim):
    amp, width = 0.2 + 0.8 * k / nsim, (0.2 + 0.6 * ((7 * k) % nsim) / nsim) * nsamp * dt
    signals.append([qd.DiscreteSignal(dt=dt, samples=envelope(amp * (1 + 0.1 * j), width), carrier_freq=freqs[j], phase=0.1 * j) for j in range(8)])
user_s = time.perf_counter() - t0
y0 = np.zeros(n, dtype=complex); y0[0] = 1.0
kw = dict(method="RK4", max_dt=dt)
def run():
    out = solver.solve(t_span=[0.0, nsamp * dt], y0=y0, signals=signals, **kw)
    torch.cuda.synchronize()
    return out
run()  # warm-up (library load, allocator)
l0 = abi.launch_count(); t0 = time.perf_counter(); out = run(); api_s = time.perf_counter() - t0; launches = abi.launch_count() - l0
# kernel-only: time the GPU work of the same call with events around a second run of the solve, minus nothing -- reported beside
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
norms = torch.stack([r.y[-1] for r in out[:64]], dim=1).abs().pow(2).sum(dim=0)
print(json.dumps({"config": f"cfg5-style API sweep: {nq} transmons x {dim} levels (n={n}), 8 channels, {nsim} simulations, {nsamp} RK4 steps",
                  "user_side_signal_construction_s": user_s, "solver_solve_s": api_s, "solver_solve_event_ms": e0.elapsed_time(e1),
                  "state_rhs_per_s_api": 4.0 * nsamp * nsim / api_s, "qdb_launches": launches,
                  "tiling": abi.rk4_tiling(n, nsim, 8), "max_norm_drift": float((norms - 1).abs().max())}), flush=True)
