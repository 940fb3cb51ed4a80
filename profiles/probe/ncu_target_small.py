"""ncu target: sweep-mode fused RK4 at cfg2 (n=32, K=8, B=1024)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qiskit_dynamics_b200 as qd
from qiskit_dynamics_b200 import _abi as abi
from qiskit_dynamics_b200.solvers import stage_time_grid
from oracle import numpy_oracle as orc
n, K, B, S = 32, 8, 1024, int(os.environ.get("QDB_S", "20"))
H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 2002)
m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=[qd.Signal(*s) for s in sig], rotating_frame=H0)
coll = m._collection(); ops_p, stat_p = coll.packed(); mu = m._frame_freqs()
times = stage_time_grid(0.0, 1e-3, S)
y = m.rotating_frame.state_into_frame_basis(qd.asarray(Y))
base = m._signal_table(times); amp = 0.5 + np.arange(B) / B
coeff = torch.from_numpy(np.ascontiguousarray(base[:, :, None] * amp[None, None, :])).cuda()
for _ in range(3):
    abi.rk4_steps(n, coll.operators, coll.static_operator, ops_p, stat_p, coeff, mu, times, 1e-3, y, S, per_col=True)
torch.cuda.synchronize()
print("done")
