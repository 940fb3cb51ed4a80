"""rk4_ozaki_kernel (int8 tensor-core emulation of the fp64 contraction) against rk4_shared3m_kernel on the same generator
table: max column-L2 difference and time per RK4 step.  python profiles/probe/ozaki_probe.py n B S"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi
n, B, S = (int(x) for x in sys.argv[1:4])
rng = np.random.default_rng(5)
A = rng.standard_normal((2 * S + 1, n, n)) + 1j * rng.standard_normal((2 * S + 1, n, n))
table = torch.from_numpy((A - A.conj().transpose(0, 2, 1)) * (5.0 / np.sqrt(2 * n))).cuda().contiguous()
y0 = torch.from_numpy(rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))).cuda()
y0 = y0 / torch.linalg.vector_norm(y0, dim=0, keepdim=True)
h = 1e-3
packed3 = abi.to_packed3m(abi.pack_operators(table))
rowmajor = table.reshape(2 * S + 1, n * n).contiguous()
ws = torch.empty(int(abi.lib().qdb_rk4_ozaki_workspace_bytes(S)), dtype=torch.uint8, device="cuda")
def run_ref():
    y = y0.clone(); abi.rk4_table_steps(n, packed3, h, y, S, layout=abi.LAYOUT_PACKED3M); return y
def run_oz():
    y = y0.clone(); abi.rk4_ozaki_steps(n, rowmajor, h, y, S, workspace=ws); return y
yr, yo = run_ref(), run_oz()
torch.cuda.synchronize()
err = float(torch.linalg.vector_norm(yo - yr, dim=0).max())
drift = float((torch.linalg.vector_norm(yo, dim=0) - 1).abs().max())
def timeit(fn):
    best = 1e30
    for it in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 2: best = min(best, e0.elapsed_time(e1))
    return best
t_ref, t_oz = timeit(run_ref), timeit(run_oz)
print(json.dumps({"n": n, "B": B, "S": S, "max_col_l2_ozaki_vs_dmma": err, "unitarity_drift_ozaki": drift,
                  "dmma_us_per_step": t_ref * 1e3 / S, "ozaki_us_per_step_incl_slicing": t_oz * 1e3 / S, "speedup": t_ref / t_oz}))

import ctypes
buf = (ctypes.c_longlong * 64)()
try:
    abi.lib().qdb_ozaki_debug(buf)
    t = list(buf)
    base = t[0]
    print(json.dumps({"mma_wait_a": t[1] - t[0], "mma_wait_b": t[2] - t[1], "mma_issue_and_drain_waits": t[3] - t[2],
                      "epi_stage_start_rel": t[8] - base, "epi_group_done_rel_g_descending": [t[8 + g] - base for g in range(7, 1, -1) if t[8 + g]],
                      "epi_combine_done_rel": t[20] - base, "epi_combine_done_warp15_rel": t[25] - base, "slice_before_bar_rel": t[22] - base, "slice_after_bar_rel": t[23] - base, "slice_stored_rel": t[24] - base, "per_warp_after_bar": [t[48 + w] - base for w in range(16)], "per_warp_stored": [t[32 + w] - base for w in range(16)], "epi_slice_done_rel": t[21] - base, "mma_stage_end_rel": t[3] - base}))
except Exception as ex:
    print("no debug", ex)
