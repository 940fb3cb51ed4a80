"""Times the per-column-signal RK4 kernels (qdb_rk4_steps_c128, sig_mode 1) at n, K, B, S from the command line under the
tiling overrides in the environment (QDB_FORCE_SWEEPF="WR,WC,MR", QDB_FORCE_SWEEPF_NCW).  One JSON line."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_workloads as W
import qiskit_dynamics_b200 as qd
from qiskit_dynamics_b200 import _abi as abi
from qiskit_dynamics_b200.solvers import stage_time_grid
n, K, B, S = (int(x) for x in sys.argv[1:5])
H0, Hs, Y, sig = W.schrodinger(n, K, 1, 2005)
m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=[qd.Signal(*s) for s in sig], rotating_frame=H0)
c = m._collection(); p_ops, p_stat = c.packed()
t = stage_time_grid(0.0, 1e-3, S)
base = torch.from_numpy(m._signal_table(t)).cuda()
amp = 0.5 + torch.arange(B, dtype=torch.float64, device="cuda") / B
coeff = (base[:, :, None] * amp[None, None, :]).contiguous()
y0 = m.rotating_frame.state_into_frame_basis(qd.asarray(np.repeat(Y, B, axis=1))); y = y0.clone()
def run():
    y.copy_(y0); abi.rk4_steps(n, c.operators, c.static_operator, p_ops, p_stat, coeff, m._frame_freqs(), t, 1e-3, y, S, per_col=True)
best = 1e30
for it in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    if it >= 2: best = min(best, e0.elapsed_time(e1))
alg = S * B * (4 * ((4 * K + 8) * n * n + 12 * n) + 28 * n)
print(json.dumps({"n": n, "K": K, "B": B, "S": S, "us_per_step": best * 1e3 / S, "alg_tflops": alg / best * 1e-9,
                  "drift": float((torch.linalg.vector_norm(y, dim=0) - 1).abs().max()), "tiling": abi.rk4_tiling(n, B, K),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("QDB_")}}))
