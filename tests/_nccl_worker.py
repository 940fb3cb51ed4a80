"""Worker for test_nccl_sharded_sweep (torch.distributed.run, nccl backend, one rank per GPU): a cfg5-style parameter
sweep -- two three-level transmons, 4 drive/control channels with sampled Gaussian-square envelopes whose amplitude and
width vary per simulation -- split over the ranks by distributed.solver_solve_sharded, with ONE gather of the
memory-slot outcome probabilities, checked against the same sweep solved on a single GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qiskit_dynamics_b200 as qd  # noqa: E402
from qiskit_dynamics_b200 import distributed as D  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, w = D.world()
qd.set_default_device(f"cuda:{local_rank}")

dim = 3
a = np.diag(np.sqrt(np.arange(1, dim)), 1).astype(complex)
N = a.conj().T @ a
I = np.eye(dim, dtype=complex)
w0, w1, alpha, J = 2 * np.pi * 5.0, 2 * np.pi * 5.1, 2 * np.pi * -0.33, 2 * np.pi * 0.002
H0 = (w0 * np.kron(I, N) + 0.5 * alpha * np.kron(I, N @ (N - I)) + w1 * np.kron(N, I) + 0.5 * alpha * np.kron(N @ (N - I), I)
      + J * (np.kron(a, a.conj().T) + np.kron(a.conj().T, a)))
drive0, drive1 = 2 * np.pi * 0.02 * np.kron(I, a + a.conj().T), 2 * np.pi * 0.02 * np.kron(a + a.conj().T, I)
ops = [drive0, drive1, drive0, drive1]  # d0, d1, u0 (qubit 0 at qubit 1's frequency), u1
solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=ops, rotating_frame=H0)
dt, nsamp, nsim = 0.222, 24, 11 * w + 3  # ragged split
freqs = [w0 / (2 * np.pi), w1 / (2 * np.pi), w1 / (2 * np.pi), w0 / (2 * np.pi)]


def envelope(amp, width):
    t = (np.arange(nsamp) + 0.5) * dt
    c, rise = t[-1] / 2 + dt / 2, 0.15 * nsamp * dt
    flat = np.clip((np.abs(t - c) - width / 2) / rise, 0.0, None)
    return amp * np.exp(-0.5 * flat**2).astype(complex)


signals = []
for k in range(nsim):
    amp, width = 0.2 + 0.8 * k / nsim, (0.2 + 0.6 * ((7 * k) % nsim) / nsim) * nsamp * dt
    signals.append([qd.DiscreteSignal(dt=dt, samples=envelope(amp * (1 + 0.1 * j), width), carrier_freq=freqs[j], phase=0.1 * j)
                    for j in range(4)])
y0 = np.zeros(dim * dim, dtype=complex)
y0[0] = 1.0
meas = qd.FinalStateMeasurement(solver.model, subsystem_dims=[dim, dim], measurement_subsystems=[0, 1], max_outcome_level=1)
kw = dict(method="RK4", max_dt=dt / 4)
t_span = [0.0, nsamp * dt]
before = qd._abi.launch_count()
local, probs = D.solver_solve_sharded(solver, t_span, y0, signals, measurement=meas, **kw)
assert qd._abi.launch_count() > before, "no libqdb kernel ran on this rank"
lo, hi = D.shard_bounds(nsim)
assert len(local) == hi - lo and probs.shape == (len(meas.labels), nsim)
# single-GPU answer (every rank solves the whole list itself)
full = solver.solve(t_span=t_span, y0=y0, signals=signals, **kw)
ref = meas.probabilities(t_span[1], torch.stack([r.y[-1] for r in full], dim=-1))
err = float((probs - ref).abs().max())
assert err < 1e-12, err
assert float((probs.sum(dim=0) - 1).abs().max()) < 1e-12
spread = float((ref.max(dim=1).values - ref.min(dim=1).values).max())
assert spread > 1e-3, "the sweep points must differ"
# states instead of observables
_, states = D.solver_solve_sharded(solver, t_span, y0, signals, **kw)
assert states.shape == (dim * dim, nsim)
assert float((states - torch.stack([r.y[-1] for r in full], dim=-1)).abs().max()) < 1e-12
dist.barrier()
print(f"NCCL_OK rank={rank} world={w} sims={nsim} local={hi - lo} max|dP|={err:.1e} spread={spread:.2e}", flush=True)
dist.destroy_process_group()
