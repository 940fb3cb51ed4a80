"""Tile-configuration experiment for the fused shared-signal RK4 kernel (QDB_FORCE_CFG hook)."""
import sys, os, json, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from qiskit_dynamics_b200 import _abi as abi
    n, B, S = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    table = (torch.randn(2 * S + 1, abi.packed_elems(n), dtype=torch.complex128, device="cuda") * 0.01)
    y = torch.randn(n, B, dtype=torch.complex128, device="cuda")
    for _ in range(2): abi.rk4_table_steps(n, table, 1e-3, y, S)
    torch.cuda.synchronize(); best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); abi.rk4_table_steps(n, table, 1e-3, y, S); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = S * B * (4 * (8 * n * n + 12 * n) + 28 * n)
    print(json.dumps({"cfg": os.environ.get("QDB_FORCE_CFG", "auto"), "n": n, "B": B, "us_per_step": best * 1e3 / S, "tflops": flops / best * 1e-9}))
else:
    for n, B in ((128, 4096), (128, 4736), (128, 512), (64, 4096), (32, 1024)):
        for cfg in ("auto", "8,1,2,4", "4,1,4,2", "4,2,4,2", "8,1,2,2", "4,1,4,1", "8,1,1,4", "4,1,2,4", "4,2,2,4", "4,1,1,4", "4,1,1,2", "4,2,1,4"):
            env = dict(os.environ)
            if cfg != "auto": env["QDB_FORCE_CFG"] = cfg
            r = subprocess.run([sys.executable, __file__, "child", str(n), str(B), "100"], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or ("FAIL " + cfg + " " + r.stderr.strip()[-200:]), flush=True)
