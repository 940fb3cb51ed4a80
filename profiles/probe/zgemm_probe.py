"""zgemm: 3-product kernel vs 4-product kernel (QDB_ZGEMM_4M=1) on the shapes the hot path uses, CUDA-event timing
(best of 5 after 2 warm-ups; operands larger than nothing in particular -- L2 resident for the small shapes, as in use),
against the live DMMA probe.  One JSON line per (shape, kernel).

    python profiles/probe/zgemm_probe.py
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from qiskit_dynamics_b200 import _abi as abi

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

peak = abi.dmma_probe()
shapes = [("propagator product (cfg3 expm)", 729, 729, 729), ("propagator apply (cfg3)", 729, 4096, 729),
          ("one batched RHS (cfg4)", 128, 4096, 128), ("generic RK4 stage n=264", 264, 4096, 264),
          ("square 2048", 2048, 2048, 2048), ("square 4096", 4096, 4096, 4096)]
for name, M, N, K in shapes:
    A = torch.randn(M, K, dtype=torch.complex128, device="cuda"); Bm = torch.randn(K, N, dtype=torch.complex128, device="cuda")
    C = torch.empty(M, N, dtype=torch.complex128, device="cuda")
    row = {"shape": name, "M": M, "N": N, "K": K, "dmma_peak_tflops": peak}
    for tag, env in (("m3", None), ("m4", "1")):
        if env: os.environ["QDB_ZGEMM_4M"] = env
        else: os.environ.pop("QDB_ZGEMM_4M", None)
        ms = timeit(lambda: abi.zgemm(A, Bm, out=C))
        row[tag + "_us"] = ms * 1e3; row[tag + "_alg_tflops"] = 8.0 * M * N * K / ms * 1e-9; row[tag + "_alg_frac"] = row[tag + "_alg_tflops"] / peak
    os.environ.pop("QDB_ZGEMM_4M", None)
    ms = timeit(lambda: torch.matmul(A, Bm, out=C))
    row["cublas_us"] = ms * 1e3; row["cublas_tflops"] = 8.0 * M * N * K / ms * 1e-9
    row["speedup_m3_over_m4"] = row["m4_us"] / row["m3_us"]
    print(json.dumps(row), flush=True)
