"""ncu target: the headline launch -- generator table + rk4_shared_kernel at cfg4 (n=128, K=8, B=4096).
QDB_S = RK4 steps per launch (default 10; bench.py uses 100)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi  # noqa: E402

n, B, S = int(os.environ.get("QDB_N", "128")), int(os.environ.get("QDB_B", "4096")), int(os.environ.get("QDB_S", "10"))
torch.manual_seed(0)
table = torch.randn(2 * S + 1, abi.packed_elems(n), dtype=torch.complex128, device="cuda") * 0.05
layout = abi.LAYOUT_PACKED if os.environ.get("QDB_NO_3M") == "1" else abi.rk4_table_layout(n, B)
if layout == abi.LAYOUT_PACKED3M:
    table = abi.to_packed3m(table)
y = torch.randn(n, B, dtype=torch.complex128, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for _ in range(3):
    flush.zero_()
    abi.rk4_table_steps(n, table, 1e-3, y, S, layout=layout)
torch.cuda.synchronize()
print("tiling", abi.rk4_tiling(n, B))
