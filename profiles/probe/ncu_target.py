"""Short ncu target: one generator-table launch + one fused RK4 launch (S steps) at the headline shape."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qiskit_dynamics_b200 as qd
from qiskit_dynamics_b200 import _abi as abi
from qiskit_dynamics_b200.solvers import stage_time_grid
from oracle import numpy_oracle as orc
n, K, B = 128, 8, int(os.environ.get("QDB_B", "4096"))
S = int(os.environ.get("QDB_S", "10"))
mode = os.environ.get("QDB_MODE", "shared")
H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 2004)
m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=[qd.Signal(*s) for s in sig], rotating_frame=H0)
coll = m._collection(); ops_p, stat_p = coll.packed(); mu = m._frame_freqs()
times = stage_time_grid(0.0, 1e-3, S)
y = m.rotating_frame.state_into_frame_basis(qd.asarray(Y))
if mode == "shared":
    coeff = torch.from_numpy(m._signal_table(times)).cuda(); td = torch.from_numpy(times).cuda()
    for _ in range(int(os.environ.get("QDB_REPS", "3"))):
        table = abi.generator(n, ops_p, stat_p, coeff, mu, td, layout=abi.LAYOUT_PACKED)
        abi.rk4_table_steps(n, table, 1e-3, y, S)
else:
    base = m._signal_table(times); amp = 0.5 + np.arange(B) / B
    coeff = torch.from_numpy(np.ascontiguousarray(base[:, :, None] * amp[None, None, :])).cuda()
    for _ in range(int(os.environ.get("QDB_REPS", "3"))):
        abi.rk4_steps(n, coll.operators, coll.static_operator, ops_p, stat_p, coeff, mu, times, 1e-3, y, S, per_col=True)
torch.cuda.synchronize()
print("done", float(torch.linalg.vector_norm(y, dim=0).mean()))
