"""Times rk4_shared_kernel alone with and without split (2-CTA cluster) tiling."""
import json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from qiskit_dynamics_b200 import _abi as abi
    n, B, S = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    table = torch.randn(2 * S + 1, abi.packed_elems(n), dtype=torch.complex128, device="cuda") * 0.01
    y = torch.randn(n, B, dtype=torch.complex128, device="cuda")
    for _ in range(2):
        abi.rk4_table_steps(n, table, 1e-3, y, S)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); abi.rk4_table_steps(n, table, 1e-3, y, S); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = S * B * (4 * (8 * n * n + 12 * n) + 28 * n)
    print(json.dumps({"nosplit": os.environ.get("QDB_NO_SPLIT", "0"), "n": n, "B": B, "tiling": abi.rk4_tiling(n, B),
                      "us_per_step": best * 1e3 / S, "tflops": flops / best * 1e-9}))
else:
    for n, B in ((128, 4096), (128, 4144), (128, 8192), (128, 2048), (100, 4096), (96, 4096)):
        for ns in ("0", "1"):
            env = dict(os.environ, QDB_NO_SPLIT=ns)
            r = subprocess.run([sys.executable, __file__, "child", str(n), str(B), "100"], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or ("FAIL " + r.stderr.strip()[-300:]), flush=True)
