"""Complex GEMM through qdb_zgemm_c128 (QDB_ZGEMM_INT8=0: fp64 DMMA kernels, =2: int8 tensor-core emulation forced) against
torch.matmul (cuBLAS ZGEMM): time and normwise error.  python profiles/probe/zgemm_int8_probe.py [M N K ...]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi
shapes = [(729, 4096, 729), (729, 729, 729), (264, 4096, 264), (128, 4096, 128), (2048, 2048, 2048), (300, 1000, 200), (100, 37, 129)]
if len(sys.argv) > 3:
    v = [int(x) for x in sys.argv[1:]]
    shapes = [tuple(v[i:i + 3]) for i in range(0, len(v), 3)]
g = torch.Generator(device="cuda").manual_seed(1)
def bench(fn, reps=5):
    best = 1e30
    for it in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 2: best = min(best, e0.elapsed_time(e1))
    return best * 1e3
for M, N, K in shapes:
    A = torch.randn(M, K, dtype=torch.complex128, device="cuda", generator=g)
    B = torch.randn(K, N, dtype=torch.complex128, device="cuda", generator=g)
    ref = A @ B
    out = abi.zgemm(A, B)
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())
    # all epilogue options at once
    C0 = torch.randn(M, N, dtype=torch.complex128, device="cuda", generator=g)
    cs = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    pre = torch.exp(1j * torch.randn(K, dtype=torch.float64, device="cuda", generator=g))
    post = torch.exp(1j * torch.randn(M, dtype=torch.float64, device="cuda", generator=g))
    alpha, beta = 0.3 - 1.2j, -0.7 + 0.4j
    c = C0.clone()
    abi.zgemm(A, B, out=c, alpha=alpha, beta=beta, colscale=cs, pre=pre, post=post)
    ref2 = beta * C0 + alpha * cs[None, :] * post[:, None] * (A @ (pre[:, None] * B))
    err2 = float((c - ref2).abs().max() / ref2.abs().max())
    t = bench(lambda: abi.zgemm(A, B, out=out))
    tc = bench(lambda: torch.matmul(A, B, out=ref))
    print(json.dumps({"M": M, "N": N, "K": K, "mode": os.environ.get("QDB_ZGEMM_INT8", "default"), "us": t, "cublas_us": tc,
                      "speedup_vs_cublas": tc / t, "alg_tflops": 8.0 * M * N * K / t * 1e-6, "rel_err": err, "rel_err_full_epilogue": err2}))
