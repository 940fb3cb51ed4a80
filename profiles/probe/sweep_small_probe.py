"""Sweep-mode RK4 on small systems: formed-generator kernel vs the shared-memory-resident operator-pass kernel
(QDB_SWEEP_KERNEL=formed / legacy), us per RK4 step.  One JSON line per shape."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi
from qiskit_dynamics_b200.solvers import stage_time_grid
from oracle import numpy_oracle as orc

def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

shapes = [(32, 8, 1024), (32, 8, 4096), (32, 8, 32768), (32, 4, 1024), (32, 3, 8192), (16, 8, 1024), (16, 4, 16384), (8, 3, 8192),
          (24, 6, 2048), (4, 3, 65536), (27, 8, 4096)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for n, K, B in shapes:
    S = 100
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, 1, 7)
    Gd, G, d, U = orc.generator_model_operators(H0, Hs, H0)
    specs = [orc.SigSpec(*s) for s in sig]
    h = 1e-3; times = stage_time_grid(0.0, h, S); base = orc.signal_list_values(specs, times)
    coeff = (dev(base)[:, :, None] * dev(0.5 + np.arange(B) / B)[None, None, :]).contiguous()
    Gdv, Gdd = dev(G), dev(Gd); Gp, Gdp = abi.pack_operators(Gdv), abi.pack_operators(Gdd[None])[0]
    mu = dev(-np.imag(d)); y0 = dev(np.repeat(U.conj().T @ Y, B, axis=1)); y = y0.clone()
    row = {"n": n, "K": K, "B": B}
    outs = {}
    for kern in ("formed", "legacy"):
        os.environ["QDB_SWEEP_KERNEL"] = kern
        def run():
            y.copy_(y0); abi.rk4_steps(n, Gdv, Gdd, Gp, Gdp, coeff, mu, times, h, y, S, per_col=True)
        row[kern + "_us_per_step"] = timeit(run) * 1e3 / S
        outs[kern] = y.clone()
        row[kern + "_tiling"] = abi.rk4_tiling(n, B, K)
    row["speedup_formed"] = row["legacy_us_per_step"] / row["formed_us_per_step"]
    row["max_col_diff"] = float(torch.linalg.vector_norm(outs["formed"] - outs["legacy"], dim=0).max())
    print(json.dumps(row), flush=True)
