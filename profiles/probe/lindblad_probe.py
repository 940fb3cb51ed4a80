"""Row f2 measured: the cfg3 system (dim 27, 3 drive operators, 6 static dissipators, 4096 density matrices) stepped with
RK4 in NON-vectorised form (qdb_lindblad_rk4_steps_c128, O(n^3) per matrix) against the vectorised form (729 x 729
generator on 4096 columns: fused RK4 and the exponential stepper).  One JSON line."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_workloads as W
import qiskit_dynamics_b200 as qd
from qiskit_dynamics_b200 import _abi as abi

def timeit(fn, reps=4, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H0, Hs, Ls, Y, sig = W.cfg3(B)
n = 27
kw = dict(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=[qd.Signal(*s) for s in sig],
          static_dissipators=Ls, rotating_frame=np.diag(H0).real)
mm, mv = qd.LindbladModel(vectorized=False, **kw), qd.LindbladModel(vectorized=True, **kw)
rho = qd.asarray(np.ascontiguousarray(Y.T.reshape(B, n, n).transpose(0, 2, 1)))
yv = qd.asarray(Y)
S = 100
out = {}
r_m = qd.solve_lmde(mm, t_span=[0, S * 1e-3], y0=rho, method="RK4", max_dt=1e-3)
r_v = qd.solve_lmde(mv, t_span=[0, S * 1e-3], y0=yv, method="RK4", max_dt=1e-3)
err = float(torch.linalg.vector_norm(r_m.y[-1].permute(2, 1, 0).reshape(n * n, B) - r_v.y[-1], dim=0).max())
ms_m = timeit(lambda: qd.solve_lmde(mm, t_span=[0, S * 1e-3], y0=rho, method="RK4", max_dt=1e-3))
ms_v = timeit(lambda: qd.solve_lmde(mv, t_span=[0, S * 1e-3], y0=yv, method="RK4", max_dt=1e-3))
ms_e = timeit(lambda: qd.solve_lmde(mv, t_span=[0, 0.1], y0=yv, method="scipy_expm", max_dt=1e-2))
rhs_m = timeit(lambda: mm(0.013, rho))
rhs_v = timeit(lambda: mv(0.013, yv))
peak = abi.dmma_probe()
J = 6
alg = S * 4 * B * (2 + 2 * J) * 8 * n**3
print(json.dumps({"config": f"cfg3 system, {B} density matrices of dim {n}, {J} static dissipators", "rk4_steps": S,
                  "matrix_form_rk4_us_per_step": ms_m * 1e3 / S, "vectorized_rk4_us_per_step": ms_v * 1e3 / S,
                  "vectorized_expm_us_per_step_dt1e-2": ms_e * 1e3 / 10, "matrix_vs_vectorized_max_col_l2": err,
                  "matrix_form_alg_tflops": alg / ms_m * 1e-9, "matrix_form_alg_frac": alg / ms_m * 1e-9 / peak,
                  "rhs_call_matrix_form_us": rhs_m * 1e3, "rhs_call_vectorized_us": rhs_v * 1e3, "dmma_peak_tflops": peak}))
