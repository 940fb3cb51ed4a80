"""3M vs 4M complex product in the fused RK4 kernel: accuracy (vs 4M and vs a torch fp64 restatement) and speed."""
import json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from qiskit_dynamics_b200 import _abi as abi
    n, B, S = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    torch.manual_seed(1)
    # anti-Hermitian generators of norm ~ 10 (like -i H): the flow is unitary
    A = torch.randn(2 * S + 1, n, n, dtype=torch.complex128, device="cuda")
    G = (A - A.conj().transpose(1, 2)) * (10.0 / (2 * n) ** 0.5) * 0.5
    table = abi.pack_operators(G.contiguous())
    layout = abi.LAYOUT_PACKED3M if os.environ.get("QDB_3M") == "1" else abi.LAYOUT_PACKED
    if layout == abi.LAYOUT_PACKED3M:
        table = abi.to_packed3m(table)
    y0 = torch.randn(n, B, dtype=torch.complex128, device="cuda")
    y0 /= torch.linalg.vector_norm(y0, dim=0, keepdim=True)
    y = y0.clone()
    abi.rk4_table_steps(n, table, 1e-3, y, S, layout=layout)
    torch.cuda.synchronize()
    torch.save(y.cpu(), f"/tmp/m3_{os.environ.get('QDB_3M', '0')}.pt")
    # torch fp64 restatement on 64 columns
    yr = y0[:, :64].clone(); h = 1e-3
    for s in range(S):
        G0, G1, G2 = G[2 * s], G[2 * s + 1], G[2 * s + 2]
        k1 = G0 @ yr; k2 = G1 @ (yr + 0.5 * h * k1); k3 = G1 @ (yr + 0.5 * h * k2); k4 = G2 @ (yr + h * k3)
        yr = yr + (1.0 / 6) * h * (k1 + 2 * k2 + 2 * k3 + k4)
    err = torch.linalg.vector_norm(y[:, :64] - yr, dim=0).max().item()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); abi.rk4_table_steps(n, table, 1e-3, y, S, layout=layout); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = S * B * (4 * (8 * n * n + 12 * n) + 28 * n)
    print(json.dumps({"m3": os.environ.get("QDB_3M", "0"), "n": n, "B": B, "S": S, "err_vs_torch": err,
                      "us_per_step": best * 1e3 / S, "alg_tflops": flops / best * 1e-9}))
else:
    import torch
    shapes = ((128, 4096, 100), (128, 4096, 1000), (128, 512, 100), (64, 4096, 100), (100, 4096, 100), (200, 4096, 50), (128, 8192, 50))
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        shapes = ((128, 512, 200), (128, 256, 200), (100, 512, 200), (128, 64, 200))
    for n, B, S in shapes:
        for m3 in ("0", "1"):
            env = dict(os.environ, QDB_3M=m3)
            r = subprocess.run([sys.executable, __file__, "child", str(n), str(B), str(S)], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or ("FAIL " + r.stderr.strip()[-300:]), flush=True)
        a, b = torch.load("/tmp/m3_0.pt"), torch.load("/tmp/m3_1.pt")
        print(json.dumps({"n": n, "B": B, "S": S, "max_col_l2_3M_vs_4M": torch.linalg.vector_norm(a - b, dim=0).max().item()}), flush=True)
