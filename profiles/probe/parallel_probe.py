"""Time-parallel solvers vs the direct fused solvers at the BASELINE shapes (public API, device-resident y0, CUDA events).
One JSON line per configuration."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qiskit_dynamics_b200 as qd
from oracle import numpy_oracle as orc

def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

n, K, B, S = 128, 8, 4096, 1000
H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 2004)
m = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=[qd.Signal(*s) for s in sig], rotating_frame=H0)
y0 = qd.asarray(Y)
out = {}
for method in ("RK4", "jax_RK4_parallel"):
    res = {}
    def run(): res["y"] = qd.solve_lmde(m, t_span=[0, S * 1e-3], y0=y0, method=method, max_dt=1e-3).y[-1]
    out[method] = timeit(run); out[method + "_y"] = res["y"]
print(json.dumps({"config": "cfg4: n=128, K=8, B=4096, 1000 RK4 steps, shared signals", "direct_fused_ms": out["RK4"],
                  "time_parallel_ms": out["jax_RK4_parallel"], "speedup": out["RK4"] / out["jax_RK4_parallel"],
                  "max_col_diff": float(torch.linalg.vector_norm(out["RK4_y"] - out["jax_RK4_parallel_y"], dim=0).max())}), flush=True)

out = {}
for method in ("scipy_expm", "jax_expm_parallel"):
    res = {}
    def run(): res["y"] = qd.solve_lmde(m, t_span=[0, 0.2], y0=y0, method=method, max_dt=1e-3).y[-1]
    out[method] = timeit(run); out[method + "_y"] = res["y"]
print(json.dumps({"config": "cfg4 shape with the exponential stepper: n=128, B=4096, 200 expm steps (Magnus order 1)",
                  "direct_ms": out["scipy_expm"], "time_parallel_ms": out["jax_expm_parallel"],
                  "speedup": out["scipy_expm"] / out["jax_expm_parallel"],
                  "max_col_diff": float(torch.linalg.vector_norm(out["scipy_expm_y"] - out["jax_expm_parallel_y"], dim=0).max())}), flush=True)

n, K, B = 27, 3, 4096
H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 6, B, 2003)
mv = qd.LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=[qd.Signal(*s) for s in sig],
                      static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=True)
y0 = qd.asarray(Y)
out = {}
for method in ("scipy_expm", "jax_expm_parallel"):
    res = {}
    def run(): res["y"] = qd.solve_lmde(mv, t_span=[0, 0.2], y0=y0, method=method, max_dt=1e-2).y[-1]
    out[method] = timeit(run); out[method + "_y"] = res["y"]
print(json.dumps({"config": "cfg3: vectorised Lindblad 729, B=4096, 20 expm steps", "direct_ms": out["scipy_expm"],
                  "time_parallel_ms": out["jax_expm_parallel"], "speedup": out["scipy_expm"] / out["jax_expm_parallel"],
                  "max_col_diff": float(torch.linalg.vector_norm(out["scipy_expm_y"] - out["jax_expm_parallel_y"], dim=0).max())}), flush=True)
