"""ncu target: zgemm_kernel at M=N=K=QDB_M (default 729: the products inside the vectorised-Lindblad expm)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi
m = int(os.environ.get("QDB_M", "729")); nn = int(os.environ.get("QDB_NN", str(m)))
A = torch.randn(m, m, dtype=torch.complex128, device="cuda"); Bm = torch.randn(m, nn, dtype=torch.complex128, device="cuda")
C = torch.empty_like(Bm)
for _ in range(4): abi.zgemm(A, Bm, out=C)
torch.cuda.synchronize(); print("ok")
