"""Times the on-chip shared-signal RK4 kernel alone (qdb_rk4_table_steps_c128) for n, B, S from the command line;
also the ncu target for that kernel (QDB_NCU=1: two launches only).  One JSON line."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi
n, B, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
layout = abi.rk4_table_layout(n, B)
elems = abi.packed_elems(n) * (3 if layout == abi.LAYOUT_PACKED3M else 2) // 2
torch.manual_seed(0)
table = torch.randn(2 * S + 1, abi.packed_elems(n), dtype=torch.complex128, device="cuda") * 0.01
if layout == abi.LAYOUT_PACKED3M:
    table = abi.to_packed3m(table)
y = torch.randn(n, B, dtype=torch.complex128, device="cuda")
reps = 2 if os.environ.get("QDB_NCU") else 7
best = 1e30
for it in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); abi.rk4_table_steps(n, table, 1e-3, y, S, layout=layout); e1.record(); torch.cuda.synchronize()
    if it >= 2: best = min(best, e0.elapsed_time(e1))
flops = S * B * (4 * (8 * n * n + 12 * n) + 28 * n)
print(json.dumps({"n": n, "B": B, "S": S, "layout": layout, "tiling": abi.rk4_tiling(n, B), "us_per_step": best * 1e3 / S,
                  "alg_tflops": flops / best * 1e-9, "env": {k: v for k, v in os.environ.items() if k.startswith("QDB_")}}))
