"""Generic (n > 256) shared-signal RK4 -- one GEMM with the RK4 epilogue per stage -- on the int8 GEMM against the DMMA GEMM
(QDB_ZGEMM_INT8=0): us per RK4 step at the vectorised-Lindblad size.  python profiles/probe/generic_rk4_probe.py [n B S]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi
n, B, S = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (729, 4096, 10)
g = torch.Generator(device="cuda").manual_seed(3)
G = torch.randn(2, n, n, dtype=torch.complex128, device="cuda", generator=g) / np.sqrt(n)
Gd = torch.randn(n, n, dtype=torch.complex128, device="cuda", generator=g) * (3 / np.sqrt(n))
Y = torch.randn(n, B, dtype=torch.complex128, device="cuda", generator=g)
times = np.arange(2 * S + 1) * 5e-4
coeff = torch.from_numpy(np.stack([np.cos(3 * times), np.sin(2 * times)], axis=1)).cuda()
out = {}
for mode in ("0", "1"):
    os.environ["QDB_ZGEMM_INT8"] = mode
    best = 1e30
    for it in range(4):
        y = Y.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); abi.rk4_steps(n, G, Gd, None, None, coeff, None, times, 1e-3, y, S); e1.record(); torch.cuda.synchronize()
        if it >= 1: best = min(best, e0.elapsed_time(e1))
    out[mode] = (best * 1e3 / S, y)
err = float(torch.linalg.vector_norm(out["1"][1] - out["0"][1], dim=0).max() / torch.linalg.vector_norm(out["0"][1], dim=0).max())
print(json.dumps({"n": n, "B": B, "rk4_steps": S, "dmma_us_per_step": out["0"][0], "int8_us_per_step": out["1"][0],
                  "speedup": out["0"][0] / out["1"][0], "rel_col_l2_int8_vs_dmma": err}))
