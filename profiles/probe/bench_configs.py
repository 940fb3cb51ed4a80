"""Secondary configurations of BASELINE.json (cfg2, cfg3, cfg5-like) through the public API, device-resident inputs,
CUDA-event timing, against the live DMMA peak.  One JSON line per configuration (sweep-mode
throughput is reported beside the headline because the shared-signal shortcut does not exist there).

    python profiles/probe/bench_configs.py [cfg1] [cfg2] [cfg3] [cfg5] [cfg4sweep] [rhs]
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qiskit_dynamics_b200 as qd
from qiskit_dynamics_b200 import _abi as abi
from qiskit_dynamics_b200.solvers import stage_time_grid
from oracle import numpy_oracle as orc

def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
which = set(sys.argv[1:]) or {"cfg1", "cfg2", "cfg3", "cfg5", "cfg4sweep", "rhs"}
peak = abi.dmma_probe()

def sweep(name, n, K, B, S, seed):
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, 1, seed)
    Gd, G, d, U = orc.generator_model_operators(H0, Hs, H0)
    specs = [orc.SigSpec(*s) for s in sig]
    h = 1e-3; times = stage_time_grid(0.0, h, S); base = orc.signal_list_values(specs, times)
    amp = 0.5 + np.arange(B) / B
    coeff = (dev(base)[:, :, None] * dev(amp)[None, None, :]).contiguous()
    Gdv, Gdd = dev(G), dev(Gd); Gp, Gdp = abi.pack_operators(Gdv), abi.pack_operators(Gdd[None])[0]
    mu = dev(-np.imag(d)); y0 = dev(np.repeat(U.conj().T @ Y, B, axis=1)); y = y0.clone()
    def run():
        y.copy_(y0); abi.rk4_steps(n, Gdv, Gdd, Gp, Gdp, coeff, mu, times, h, y, S, per_col=True)
    ms = timeit(run)
    alg = S * B * (4 * ((4 * K + 8) * n * n + 12 * n) + 28 * n)
    tiling = abi.rk4_tiling(n, B, K)
    # executed flops: operator-pass kernels run K+1 complex passes (8 flops per element each); the formed-generator
    # kernel runs 2 Kpad FMAs to form an element and 4 to use it, i.e. (4 Kpad + 8) n^2 flops = the algorithmic count
    exe = S * B * 4 * ((4 * ((K + 3) // 4 * 4) + 8) if tiling["m3"] == 2 else (K + 1) * 8) * n * n
    kernel = "rk4_sweepf_kernel" if tiling["m3"] == 2 else ("rk4_sweep_small_kernel" if n <= 32 and K <= 16 else "rk4_sweep_kernel")
    print(json.dumps({"config": name, "kernel": kernel, "sweep_kernel_env": os.environ.get("QDB_SWEEP_KERNEL", "auto"), "n": n, "K": K, "B": B, "rk4_steps": S, "ms": ms,
                      "us_per_step": ms * 1e3 / S, "state_rhs_per_s": 4 * S * B / ms * 1e3, "alg_tflops": alg / ms * 1e-9,
                      "exec_tflops": exe / ms * 1e-9, "dmma_peak_tflops": peak, "alg_frac": alg / ms * 1e-9 / peak,
                      "exec_frac": exe / ms * 1e-9 / peak, "tiling": tiling,
                      "unitarity_drift": float((torch.linalg.vector_norm(y, dim=0) - 1).abs().max())}), flush=True)

if "cfg1" in which:
    # BASELINE configs[0]: 2-qubit HamiltonianModel, one Rabi drive, fixed-step RK4, T = 10, max_dt = 1e-3 (10 000 steps,
    # ONE state): latency bound by construction -- one fused launch walks all steps; the time-parallel solver builds the
    # 10 000 4 x 4 propagators in batched launches.  CPU figure: the oracle port (the reference's NumPy path) here.
    X = np.array([[0, 1], [1, 0]], dtype=complex); Z = np.diag([1.0, -1.0]).astype(complex); I2 = np.eye(2, dtype=complex)
    H0 = 2 * np.pi * 5 * (np.kron(Z, I2) + np.kron(I2, Z)) / 2; H1 = 2 * np.pi * 0.1 * np.kron(X, I2) / 2
    y0 = np.array([1.0, 0, 0, 0], dtype=complex)
    model = qd.HamiltonianModel(static_operator=H0, operators=[H1], signals=[qd.Signal(1.0, 5.0)], rotating_frame=H0)
    yd = qd.asarray(y0)
    res = {}
    def run(method):
        def f(): res[method] = qd.solve_lmde(model, t_span=[0, 10.0], y0=yd, method=method, max_dt=1e-3).y[-1]
        return f
    ms_direct = timeit(run("RK4"), reps=3, warm=1); ms_par = timeit(run("jax_RK4_parallel"), reps=3, warm=1)
    t0 = time.perf_counter()
    _, yo = orc.solve_hamiltonian(H0, [H1], [orc.SigSpec(1.0, 5.0)], H0, [0, 10.0], y0, 1e-3)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"config": "cfg1: 2-qubit HamiltonianModel, 1 Rabi drive, RK4, 10000 steps, one state", "fused_rk4_ms": ms_direct,
                      "time_parallel_ms": ms_par, "numpy_port_cpu_ms": cpu_s * 1e3, "us_per_step_fused": ms_direct * 1e3 / 10000,
                      "err_fused_vs_oracle": float(np.linalg.norm(res["RK4"].cpu().numpy() - yo[-1])),
                      "err_parallel_vs_oracle": float(np.linalg.norm(res["jax_RK4_parallel"].cpu().numpy() - yo[-1]))}), flush=True)

if "cfg2" in which: sweep("cfg2: dim-32, 8 drive operators, batch-1024 amplitude sweep, RK4", 32, 8, 1024, 500, 2002)
if "cfg5" in which: sweep("cfg5-like: dim-81 (4 three-level transmons), 8 channels, 8192 sweep points per GPU, RK4", 81, 8, 8192, 20, 2005)
if "cfg4sweep" in which: sweep("cfg4 shape in sweep mode: dim-128, K=8, batch 4096", 128, 8, 4096, 20, 2004)

if "cfg3" in which:
    n, K, B = 27, 3, 4096
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 6, B, 2003)
    model = qd.LindbladModel(static_hamiltonian=H0, hamiltonian_operators=Hs, hamiltonian_signals=[qd.Signal(*s) for s in sig],
                             static_dissipators=Ls, rotating_frame=np.diag(H0).real, vectorized=True)
    y0 = qd.asarray(Y)
    S = 4
    def run(): return qd.solve_lmde(model, t_span=[0, S * 1e-2], y0=y0, method="scipy_expm", max_dt=1e-2)
    ms = timeit(run, reps=3, warm=1)
    m = n * n
    apply_flops = S * 8 * m * m * B
    print(json.dumps({"config": "cfg3: 3-transmon vectorised Lindblad dim 27 (729), 6 collapse ops, expm stepper, batch 4096",
                      "n": m, "K": K, "B": B, "expm_steps": S, "ms": ms, "ms_per_step": ms / S, "steps_per_s": S / ms * 1e3,
                      "column_steps_per_s": S * B / ms * 1e3, "apply_tflops_lower_bound": apply_flops / ms * 1e-9,
                      "dmma_peak_tflops": peak}), flush=True)
    A = torch.randn(m, m, dtype=torch.complex128, device="cuda"); A = (A - A.conj().T) * 0.01
    Bm = torch.randn(m, B, dtype=torch.complex128, device="cuda"); C = torch.empty_like(Bm)
    ms = timeit(lambda: abi.zgemm(A, Bm, out=C))
    print(json.dumps({"kernel": "zgemm_kernel 729x4096x729", "ms": ms, "tflops": 8 * m * m * B / ms * 1e-9, "frac": 8 * m * m * B / ms * 1e-9 / peak}), flush=True)
    ms = timeit(lambda: abi.zgemm(A, A))
    print(json.dumps({"kernel": "zgemm_kernel 729^3", "ms": ms, "tflops": 8 * m ** 3 / ms * 1e-9, "frac": 8 * m ** 3 / ms * 1e-9 / peak}), flush=True)

if "rhs" in which:
    n, K, B = 128, 8, 4096
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, 2004)
    model = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=[qd.Signal(*s) for s in sig], rotating_frame=H0)
    model.in_frame_basis = True
    y = qd.asarray(Y)
    ms = timeit(lambda: model(0.3, y), reps=20)
    print(json.dumps({"config": "cfg4 single batched RHS call model(t, Y) (a4)", "ms": ms, "state_rhs_per_s": B / ms * 1e3,
                      "tflops": B * (8 * n * n + 12 * n) / ms * 1e-9, "frac": B * (8 * n * n + 12 * n) / ms * 1e-9 / peak}), flush=True)
