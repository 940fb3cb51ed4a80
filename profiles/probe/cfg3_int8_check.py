import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); 
import numpy as np, torch
import qiskit_dynamics_b200 as qd
import bench_workloads as W
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests', 'golden', 'fullsize.npz'))
H0, Hs, Ls, Y, sig = W.cfg3()
model = qd.LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=[qd.Signal(a, nu, ph) for a, nu, ph in sig],
                         static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=True)
y0 = Y
def run():
    return qd.solve_lmde(model, t_span=[0.0, 0.2], y0=y0, method="scipy_expm", max_dt=1e-2)
res = run(); torch.cuda.synchronize()
got = res.y[-1].cpu().numpy()[:, g["cfg3_cols"]]
err = float(np.max(np.linalg.norm(got - g["cfg3_y"], axis=0)))
best = 1e9
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
print(json.dumps({"mode": os.environ.get("QDB_ZGEMM_INT8", "default"), "cfg3_parity_max_col_l2": err, "ms_per_expm_step": best * 1e3 / 20}))
