"""Where the end-to-end (host in -> host out) overhead of the headline solve goes: cProfile + CUDA-synchronised timers."""
import cProfile, os, pstats, sys, time, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qiskit_dynamics_b200 as qd
from oracle import numpy_oracle as orc
n, K, B, S = 128, 8, 4096, 100
H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 2004)
model = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=[qd.Signal(a, nu, ph) for a, nu, ph in sig], rotating_frame=H0)
y0_host = torch.from_numpy(Y).pin_memory()
out_host = torch.empty((n, B), dtype=torch.complex128).pin_memory()
def step():
    res = qd.solve_lmde(model, t_span=[0.0, S * 1e-3], y0=y0_host, method="RK4", max_dt=1e-3)
    out_host.copy_(res.y[-1], non_blocking=True)
    torch.cuda.synchronize()
for _ in range(5): step()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
print("e2e ms: mean %.3f min %.3f" % (1e3 * np.mean(ts), 1e3 * np.min(ts)))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
# H2D / D2H alone
d = torch.empty((n, B), dtype=torch.complex128, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): d.copy_(y0_host, non_blocking=True)
torch.cuda.synchronize(); print("H2D 8 MiB ms", (time.perf_counter() - t0) / 20 * 1e3)
t0 = time.perf_counter()
for _ in range(20): out_host.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("D2H 8 MiB ms", (time.perf_counter() - t0) / 20 * 1e3)
