"""ncu target: rk4_ozaki_kernel (int8 tensor-core emulation) at cfg4 (n=128, B=4096) on a pre-sliced table.
QDB_S = RK4 steps per launch (default 10; bench.py uses 100)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi  # noqa: E402

n, B, S = int(os.environ.get("QDB_N", "128")), int(os.environ.get("QDB_B", "4096")), int(os.environ.get("QDB_S", "10"))
torch.manual_seed(0)
A = torch.randn(2 * S + 1, n, n, dtype=torch.complex128, device="cuda") * 0.3
table = (A - A.conj().transpose(1, 2)).reshape(2 * S + 1, n * n).contiguous()
planes = abi.rk4_ozaki_slice(n, table)
y = torch.randn(n, B, dtype=torch.complex128, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for _ in range(3):
    flush.zero_()
    abi.rk4_ozaki_steps(n, None, 1e-3, y, S, workspace=planes)
torch.cuda.synchronize()
print("done")
