"""ncu target: generic sweep-mode fused RK4 (QDB_N, QDB_K, QDB_B, QDB_S; default cfg5-like n=81, K=8, B=8192)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi
from qiskit_dynamics_b200.solvers import stage_time_grid
from oracle import numpy_oracle as orc
n, K, B, S = (int(os.environ.get(k, d)) for k, d in (("QDB_N", "81"), ("QDB_K", "8"), ("QDB_B", "8192"), ("QDB_S", "3")))
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, 1, 2005)
Gd, G, d, U = orc.generator_model_operators(H0, Hs, H0)
specs = [orc.SigSpec(*s) for s in sig]
h = 1e-3; times = stage_time_grid(0.0, h, S); base = orc.signal_list_values(specs, times)
coeff = (dev(base)[:, :, None] * dev(0.5 + np.arange(B) / B)[None, None, :]).contiguous()
Gdv, Gdd = dev(G), dev(Gd); Gp, Gdp = abi.pack_operators(Gdv), abi.pack_operators(Gdd[None])[0]
mu = dev(-np.imag(d)); y = dev(np.repeat(U.conj().T @ Y, B, axis=1))
for _ in range(3):
    abi.rk4_steps(n, Gdv, Gdd, Gp, Gdp, coeff, mu, times, h, y, S, per_col=True)
torch.cuda.synchronize()
print("tiling", abi.rk4_tiling(n, B, K))
