#!/usr/bin/env python
"""bench.py -- headline benchmark: RHS evals/sec, dim-128 Schrodinger in the rotating frame,
fp64 (complex128), batch 4096 per GPU, fixed-step RK4, T = 1.0, max_dt = 1e-3 (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            # the B200 arm
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference's NumPy path (CPU)

A "step" is one pass of the hot path over one batch: a fused solve of RK4_STEPS = 1000 fixed RK4 steps
(4 RHS evaluations each, the whole T = 1.0 of cfg4) on all B columns.  Metric = state-RHS evaluations per
second, whole job (sum over the N ranks; headline = weak scaling: B columns per GPU).

  value     : inputs already resident in HBM (operators, signal table, state); timed per step with CUDA events on the
              launching stream, L2 flushed between steps, max over ranks.
  e2e       : the same workload through the public API (`solve_lmde(model, ...)`) from a pinned HOST y0 to a pinned
              HOST final state -- signal table, H2D, basis changes, D2H and (N > 1) the single NCCL gather of final
              observables are all inside the timed region.
  roofline  : the dominant kernel against the fp64 tensor pipe: algorithmic flops per launch / its mean CUDA-event
              duration, over the live-measured DMMA peak, with a cuBLAS ZGEMM 4096^3 timed in the same run as the second
              witness of that peak (MEASURED_PEAKS.json has no fp64 entry).  At the headline shape the dominant kernel is
              rk4_ozaki_kernel -- the fp64 contraction emulated on the int8 tensor cores (tcgen05.mma kind::i8) -- so frac
              exceeds 1 against the fp64 roof; roofline.int8_pipe states the executed int8 MACs against the int8 tensor
              roof, and roofline.fp64_kernel the DMMA kernel (rk4_shared3m_kernel) it replaced, timed in the same run.
  parity    : max column-L2 error of the timed solve's final states on 32 chosen columns against the answers of the
              unmodified reference (committed fixture tests/golden/fullsize.npz).
  strong_scaling : the same solve on a TOTAL batch of 4096 (4096 / N columns per GPU), device-resident and e2e.
  cfg5      : BASELINE configs[4] at size -- 65 536 sweep points / N per GPU through distributed.solver_solve_sharded
              with one gather of the memory-slot probabilities (FinalStateMeasurement).
  cfg3      : BASELINE configs[2] -- vectorised Lindblad 729, batch 4096, scipy_expm T = 0.2 -- with its own roofline (the
              729^3 and 729 x 4096 x 729 products run on zgemm_ozaki_kernel, the int8 tensor-core GEMM).
  sweep_mode: per-column-signal RK4 kernels (cfg2, cfg5-like) device-resident, against the same roof.
  cpu_baseline : the reference itself (baseline/_ref through oracle/ref_shim.py; the NumPy port when absent) on the
              host cores, bounded sample, all BLAS threads and one thread.
"""
import argparse
import json
import os
import platform
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_workloads as W  # noqa: E402

N_DIM, K_OPS, BATCH = 128, 8, 4096
RK4_STEPS = 1000      # RK4 steps per bench step: cfg4 as stated (T = 1.0, max_dt = 1e-3; 4000 RHS evaluations per column)
MAX_DT = W.MAX_DT
SEED = 2004
REF_RK4_STEPS = 20    # bounded CPU sample per reference-arm step (same batch, 20 of the 1000 RK4 steps)
CPU_BASELINE_RK4_STEPS = 50  # cpu_baseline leg of the B200 arm (same batch)
CFG5_POINTS, CFG5_SAMPLES = 65536, 64

# identical in both arms: the driver compares the two `config` objects
CONFIG = {"workload": f"cfg4: dim-{N_DIM} Schrodinger in rotating frame, K={K_OPS} drive operators, batch {BATCH} per GPU, "
                      f"complex128, fixed-step RK4 max_dt={MAX_DT}",
          "n": N_DIM, "K": K_OPS, "batch_per_gpu": BATCH, "stepper": "RK4", "max_dt": MAX_DT, "seed": SEED}


def flops_per_column_step(n):
    """4 (8 n^2 + 12 n) + 28 n per RK4 step per column."""
    return 4 * (8 * n * n + 12 * n) + 28 * n


def workload(n, K, B, seed):
    """Synthetic cfg4-style inputs (bench_workloads.schrodinger); tests/test_host_cpu.py checks it draws the same
    numbers as the generator the parity tests use."""
    return W.schrodinger(n, K, B, seed)


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (baseline/_ref), else the oracle port
# ---------------------------------------------------------------------------------------------

def cpu_threads():
    """BLAS threads the CPU legs run with: all host cores (torchrun exports OMP_NUM_THREADS=1, so
    the limit is raised explicitly)."""
    return os.cpu_count() or 1


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def blas_info():
    try:
        from threadpoolctl import threadpool_info
        return [{k: d.get(k) for k in ("user_api", "internal_api", "version", "num_threads", "threading_layer", "architecture")}
                for d in threadpool_info() if d.get("user_api") == "blas"]
    except Exception:  # noqa: BLE001
        return None


class host_threads:
    def __init__(self, n):
        self.n = n

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
            self.ctx = threadpool_limits(limits=self.n, user_api="blas")
            self.ctx.__enter__()
        except Exception:  # noqa: BLE001
            self.ctx = None
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def reference_solver(n, K, B, rk4_steps, seed=SEED):
    """(run, kind): `run()` executes solve_lmde(method='RK4') over rk4_steps steps of the cfg4 workload on the host --
    through the UNMODIFIED reference when baseline/_ref (or /root/reference) is importable, else through the NumPy
    port in oracle/ (same BLAS calls, without the reference's Python dispatch overhead)."""
    H0, Hs, Y, sig = workload(n, K, B, seed)
    span = [0.0, rk4_steps * MAX_DT]
    try:
        import oracle.ref_shim as shim
        if not shim.reference_available():
            raise ImportError("no reference install")
        from qiskit_dynamics import Signal as RSignal, solve_lmde as r_solve_lmde
        from qiskit_dynamics.models import HamiltonianModel as RHamiltonianModel
        model = RHamiltonianModel(static_operator=H0, operators=Hs, signals=[RSignal(a, nu, ph) for a, nu, ph in sig],
                                  rotating_frame=H0)
        return (lambda: r_solve_lmde(model, t_span=span, y0=Y, method="RK4", max_dt=MAX_DT)), "reference"
    except Exception:  # noqa: BLE001
        from oracle import numpy_oracle as orc
        specs = [orc.SigSpec(a, nu, ph) for a, nu, ph in sig]
        return (lambda: orc.solve_hamiltonian(H0, Hs, specs, H0, span, Y, MAX_DT)), "port"


def cpu_reference_rate(n, K, B, steps, warmup, rk4_steps, threads, seed=SEED):
    """(state-RHS/s, seconds per step, kind) of the reference's solve_lmde(method='RK4') with `threads` BLAS threads."""
    run, kind = reference_solver(n, K, B, rk4_steps, seed)
    with host_threads(threads):
        for _ in range(warmup):
            run()
        t0 = time.perf_counter()
        for _ in range(steps):
            run()
        dt = time.perf_counter() - t0
    return 4.0 * rk4_steps * B * steps / dt, dt / steps, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    rate, sec_per_step, kind = cpu_reference_rate(N_DIM, K_OPS, BATCH, args.steps, args.warmup, REF_RK4_STEPS, threads)
    rate1, _, _ = cpu_reference_rate(N_DIM, K_OPS, BATCH, 1, 0, 4, 1)
    sample = (f"{REF_RK4_STEPS} of the {RK4_STEPS} RK4 steps ({4 * REF_RK4_STEPS} batched RHS calls) on the full n={N_DIM}, "
              f"K={K_OPS}, B={BATCH} batch per step; solve_lmde(method='RK4') of "
              + ("the unmodified reference (baseline/_ref via oracle/ref_shim.py)" if kind == "reference"
                 else "the NumPy port in oracle/ (reference install absent)"))
    line = {
        "impl": "reference", "metric": "rhs_evals_per_sec", "value": rate, "unit": "state-RHS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": rate, "unit": "state-RHS/s", "cores": threads, "kind": kind, "sample": sample,
                         "single_thread_value": rate1, "cpu_model": cpu_model(), "os_cpu_count": os.cpu_count(),
                         "blas": blas_info(), "numpy": np.__version__},
        "e2e": {"value": rate, "unit": "state-RHS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(busy), "power_w_max": float(max(power))}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------

class LazySweep:
    """The list of cfg5 simulations (one list of 8 DiscreteSignals each), built on demand: every rank materialises
    only the block distributed.shard_list hands it."""

    def __init__(self, qd, nsim, nsamp, freqs):
        self.qd, self.nsim, self.nsamp, self.freqs = qd, nsim, nsamp, freqs

    def __len__(self):
        return self.nsim

    def point(self, k):
        return [self.qd.DiscreteSignal(dt=W.CFG5_DT, samples=s, carrier_freq=float(self.freqs[j]), phase=ph)
                for j, (s, ph) in enumerate(W.cfg5_point(k, self.nsim, self.nsamp))]

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self.point(k) for k in range(*idx.indices(self.nsim))]
        return self.point(idx)


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as ge
    ge.build()
    import qiskit_dynamics_b200 as qd
    from qiskit_dynamics_b200 import _abi as abi
    from qiskit_dynamics_b200 import distributed as D
    from qiskit_dynamics_b200.solvers import stage_time_grid

    qd.set_default_device(f"cuda:{local_rank}")
    dev = torch.device("cuda", local_rank)
    n, K, B, S = N_DIM, K_OPS, BATCH, RK4_STEPS
    warm = max(args.warmup, 3)
    t_span = [0.0, S * MAX_DT]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2

    def make_case(Bc, seed):
        """Model + device-resident operands + pinned host buffers of one shared-signal cfg4-shaped case."""
        H0, Hs, Y, sig = workload(n, K, Bc, seed)
        model = qd.HamiltonianModel(static_operator=H0, operators=Hs, signals=[qd.Signal(a, nu, ph) for a, nu, ph in sig],
                                    rotating_frame=H0)
        coll = model._collection()
        ops_p, stat_p = coll.packed()
        mu = model._frame_freqs()
        times = stage_time_grid(0.0, MAX_DT, S)
        coeff = torch.from_numpy(model._signal_table(times)).to(dev)
        y_fb = model.rotating_frame.state_into_frame_basis(qd.asarray(Y))
        ws = torch.empty(min(abi.workspace_bytes(abi.WS_RK4, n, K, Bc, S), 1 << 30), dtype=torch.uint8, device=dev)
        case = dict(H0=H0, Hs=Hs, Y=Y, sig=sig, model=model, coll=coll, ops_p=ops_p, stat_p=stat_p, mu=mu, times=times,
                    coeff=coeff, y_fb=y_fb, y_work=y_fb.clone(), ws=ws, B=Bc,
                    y0_host=torch.from_numpy(Y).pin_memory(),
                    out_host=torch.empty((n, Bc), dtype=torch.complex128).pin_memory(),
                    obs_host=torch.empty(Bc * world, dtype=torch.float64).pin_memory(),
                    obs_all=torch.empty(Bc * world, dtype=torch.float64, device=dev),
                    zero_map=torch.zeros(n, dtype=torch.int32, device=dev))
        return case

    def device_step(c):
        """One bench step, device-resident: generator tables (chunked to the 1 GiB workspace) + fused RK4 launches."""
        c["y_work"].copy_(c["y_fb"])
        abi.rk4_steps(n, c["coll"].operators, c["coll"].static_operator, c["ops_p"], c["stat_p"], c["coeff"], c["mu"],
                      c["times"], MAX_DT, c["y_work"], S, per_col=False, workspace=c["ws"])

    def e2e_step(c):
        """Public API, pinned host in -> pinned host out; the per-column observable (squared norm, one
        outcome_prob_kernel launch) is the object of the single collective."""
        res = qd.solve_lmde(c["model"], t_span=t_span, y0=c["y0_host"], method="RK4", max_dt=MAX_DT)
        yf = res.y[-1]
        c["out_host"].copy_(yf, non_blocking=True)
        obs = abi.outcome_probabilities(yf, c["zero_map"], 1, normalize=False).reshape(-1)
        if world > 1:
            dist.all_gather_into_tensor(c["obs_all"], obs)
            c["obs_host"].copy_(c["obs_all"], non_blocking=True)
        else:
            c["obs_host"].copy_(obs, non_blocking=True)
        torch.cuda.synchronize()

    def time_device(c, steps):
        for _ in range(warm):
            device_step(c)
        barrier()
        l0 = abi.launch_count()
        ms = []
        barrier()
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            device_step(c)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        barrier()
        return max_over_ranks(float(np.mean(ms))), abi.launch_count() - l0

    def time_e2e(c, steps):
        for _ in range(warm):
            e2e_step(c)
        barrier()
        secs = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_step(c)
            secs.append(time.perf_counter() - t0)
        barrier()
        return max_over_ranks(float(np.mean(secs)) * 1e3)

    # ---------------- headline: weak scaling, B columns per GPU ----------------
    case = make_case(B, SEED + rank)  # every rank owns its own B columns
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, launches = time_device(case, args.steps)
    e2e_ms = time_e2e(case, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = 4.0 * S * B * world / (dev_ms * 1e-3)
    e2e_value = 4.0 * S * B * world / (e2e_ms * 1e-3)

    # the dominant kernel alone (one launch = one chunk of steps from a prebuilt table), CUDA events on its stream:
    # the fp64 DMMA stepper always, and the int8 tensor-core stepper when the solve above took it
    layout = abi.rk4_table_layout(n, B)
    int8_path = abi.rk4_int8_preferred(n, B)
    S_k = 100
    entry_elems = abi.packed_elems(n) * (3 if layout == abi.LAYOUT_PACKED3M else 2) // 2
    table = torch.empty((2 * S_k + 1, entry_elems), dtype=torch.complex128, device=dev)
    times_k = torch.from_numpy(case["times"][: 2 * S_k + 1]).to(dev)
    abi.generator(n, case["ops_p"], case["stat_p"], case["coeff"][: 2 * S_k + 1].contiguous(), case["mu"], times_k,
                  layout=layout, out=table)

    def time_kernel(launch):
        ms = []
        for it in range(3 + 10):
            case["y_work"].copy_(case["y_fb"])
            flush.zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            launch()
            k1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ms.append(k0.elapsed_time(k1))
        return float(np.mean(ms))

    dmma_kern_ms = time_kernel(lambda: abi.rk4_table_steps(n, table, MAX_DT, case["y_work"], S_k, layout=layout))
    del table
    kern_mean_ms = dmma_kern_ms
    if int8_path:
        tpk = abi.generator(n, case["ops_p"], case["stat_p"], case["coeff"][: 2 * S_k + 1].contiguous(), case["mu"], times_k,
                            layout=abi.LAYOUT_PACKED)
        planes = abi.rk4_ozaki_slice(n, tpk, layout=abi.LAYOUT_PACKED)
        del tpk
        kern_mean_ms = time_kernel(lambda: abi.rk4_ozaki_steps(n, None, MAX_DT, case["y_work"], S_k, workspace=planes))
        del planes

    # ---------------- strong scaling: 4096 columns in TOTAL ----------------
    if world == 1:
        strong = {"total_batch": BATCH, "batch_per_gpu": BATCH, "ms_per_step": dev_ms, "value": value, "e2e_ms_per_step": e2e_ms,
                  "e2e_value": e2e_value, "note": "N = 1: identical to the headline"}
    else:
        lo, hi = D.shard_bounds(BATCH, rank, world)
        scase = make_case(hi - lo, SEED)  # the SAME 4096-column problem on every N: rank r owns columns [lo, hi)
        H0_, Hs_, Yfull, _ = workload(n, K, BATCH, SEED)
        scase["Y"] = Yfull[:, lo:hi].copy()
        scase["y0_host"] = torch.from_numpy(scase["Y"]).pin_memory()
        scase["y_fb"] = scase["model"].rotating_frame.state_into_frame_basis(qd.asarray(scase["Y"]))
        ssteps = max(3, min(args.steps, 10))
        s_ms, _ = time_device(scase, ssteps)
        s_e2e = time_e2e(scase, ssteps)
        strong = {"total_batch": BATCH, "batch_per_gpu": hi - lo, "ms_per_step": s_ms, "value": 4.0 * S * BATCH / (s_ms * 1e-3),
                  "e2e_ms_per_step": s_e2e, "e2e_value": 4.0 * S * BATCH / (s_e2e * 1e-3), "steps": ssteps,
                  "kernel": ("rk4_ozaki_kernel (int8 tensor cores)" if abi.rk4_int8_preferred(n, hi - lo) else "fp64 DMMA, tiling below"),
                  "tiling": abi.rk4_tiling(n, hi - lo),
                  "note": "same 1000-step solve, total batch fixed at 4096 columns, split over the ranks; value = whole-job "
                          "state-RHS/s; strong-scaling efficiency = value(N) / (N * value(1))"}
        del scase

    # ---------------- cfg5 at size: 65 536 sweep points over the ranks, one gather of memory-slot probabilities ----------------
    def cfg5_record():
        H0c, opsc, freqs = W.cfg5_system()
        nc = H0c.shape[0]
        solver = qd.Solver(static_hamiltonian=H0c, hamiltonian_operators=opsc, rotating_frame=H0c)
        dims, msub, mslots = W.cfg5_measurement()
        meas = qd.FinalStateMeasurement(solver.model, subsystem_dims=dims, measurement_subsystems=msub,
                                        memory_slot_indices=mslots, max_outcome_level=1)
        y0 = np.zeros(nc, dtype=complex)
        y0[0] = 1.0
        sweep = LazySweep(qd, CFG5_POINTS, CFG5_SAMPLES, freqs)
        lo, hi = D.shard_bounds(CFG5_POINTS, rank, world)
        tb = time.perf_counter()
        local = sweep[lo:hi]  # user-side construction of this rank's signal objects (not timed: the API's own cost)
        build_s = time.perf_counter() - tb

        class Block:  # solver_solve_sharded slices the full list of simulations; hand it this rank's prebuilt block
            def __len__(self):
                return CFG5_POINTS

            def __getitem__(self, idx):
                if isinstance(idx, slice):
                    a, b, _ = idx.indices(CFG5_POINTS)
                    assert (a, b) == (lo, hi)
                    return local
                return local[idx - lo]

        kw = dict(method="RK4", max_dt=W.CFG5_DT)
        span = [0.0, CFG5_SAMPLES * W.CFG5_DT]
        times_s = []
        probs = None
        for it in range(3):
            barrier()
            t0 = time.perf_counter()
            _, probs = D.solver_solve_sharded(solver, span, y0, Block(), measurement=meas, **kw)
            torch.cuda.synchronize()
            times_s.append(time.perf_counter() - t0)
        barrier()
        sec = max_over_ranks(min(times_s[1:]))
        rec = {"points": CFG5_POINTS, "points_per_gpu": hi - lo, "n": nc, "K": 8, "rk4_steps": CFG5_SAMPLES,
               "seconds": sec, "state_rhs_per_s": 4.0 * CFG5_SAMPLES * CFG5_POINTS / sec,
               "outcomes": len(meas.labels), "gathered_shape": list(probs.shape),
               "prob_sum_max_dev": float((probs.sum(dim=0) - 1).abs().max()),
               "signal_objects_build_s_per_rank": build_s, "tiling": abi.rk4_tiling(nc, hi - lo, 8),
               "note": "Solver.solve on lists of DiscreteSignal objects (Gaussian-square, amplitude and width swept) via "
                       "distributed.solver_solve_sharded + FinalStateMeasurement: host compile of the signal program, "
                       "device signal table, sweep-mode RK4, post-processing and the one gather are inside `seconds` "
                       "(wall clock, max over ranks, best of 2 after 1 warm-up)"}
        if rank == 0:  # parity of the gathered table on 8 sweep points against the UNMODIFIED reference's answers (fixture)
            fx = np.load(os.path.join(ROOT, "tests", "golden", "fullsize.npz"))
            Pg = probs.cpu().numpy()
            rows = [meas.labels.index(str(lab)) for lab in fx["cfg5_big_labels"]]
            rec["parity_max_abs_prob_err"] = float(np.max(np.abs(Pg[np.ix_(rows, fx["cfg5_big_points"])] - fx["cfg5_big_probs"])))
            rec["parity_against"] = "tests/golden/fullsize.npz: memory-slot probabilities of 8 of the 65 536 points from the unmodified reference"
        return rec

    cfg5 = cfg5_record()

    # ---------------- rank 0: roofline, cfg3, sweep-mode kernels, parity, CPU baseline ----------------
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak_tf = abi.dmma_probe()
    # second witness of the fp64 tensor peak: cuBLAS ZGEMM 4096^3 (8 N^3 flops) in the same run
    za = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    zb = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    zbest = float("inf")
    for it in range(4):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        torch.matmul(za, zb)
        a1.record()
        torch.cuda.synchronize()
        if it >= 1:
            zbest = min(zbest, a0.elapsed_time(a1))
    cublas_tf = 8.0 * 4096**3 / zbest * 1e-9
    del za, zb

    flops_launch = float(S_k) * B * flops_per_column_step(n)
    achieved_tf = flops_launch / (kern_mean_ms * 1e-3) * 1e-12
    # DRAM traffic per launch from the committed ncu --set full capture (profiles/rk4_shared_traffic.json): the state
    # read (fixed) plus the generator-table entries, each read from HBM once; scaled to this launch's 2 S_k + 1 entries
    traffic = None
    prof = os.path.join(ROOT, "profiles", "rk4_shared_traffic.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            traffic = float(pj["dram_bytes_fixed"]) + float(pj["dram_bytes_per_table_entry"]) * (2 * S_k + 1)
        except Exception:  # noqa: BLE001
            traffic = None
    tiling = abi.rk4_tiling(n, B)
    dmma_name = (f"{'rk4_shared3m_kernel' if tiling['m3'] else 'rk4_shared_kernel'}<{tiling['row_tiles_per_warp']},"
                 f"{tiling['col_tiles_per_warp']},{'split' if tiling['split'] else 'whole'}>")
    dmma_tf = flops_launch / (dmma_kern_ms * 1e-3) * 1e-12
    fp64_kernel = {"kernel": dmma_name, "tiling": tiling, "kernel_ms": dmma_kern_ms, "achieved": dmma_tf, "frac": dmma_tf / peak_tf,
                   "pipe_frac": dmma_tf * (0.75 if tiling["m3"] else 1.0) / peak_tf, "traffic": traffic,
                   "note": "the fp64 DMMA stepper on the same table (QDB_RK4_INT8=0 makes it the solve's kernel); the "
                           "3-product kernel issues 6 flops per complex multiply-add, pipe_frac = executed DMMA flops / peak"}
    kernel_name = dmma_name
    int8_pipe = None
    if int8_path:
        # rk4_ozaki_kernel: per column tile of 32 and RHS evaluation 15 slice pairs x 4 k-steps x 2 MMAs of
        # M128 x N64 x K32 int8 (A from TMEM).  Roof: the measured TMEM-operand issue rate of 8192 MAC/clk/SM
        # (profiles/r02_m_umma_i8_probe.jsonl) x 148 SMs x the SM clock under load.
        kernel_name = "rk4_ozaki_kernel<5 byte slices, 32 columns per CTA>"
        ctas = (B + 31) // 32
        macs = float(S_k) * 4 * ctas * 15 * 4 * 2 * (128 * 64 * 32)
        sm_mhz = float((clocks or {}).get("sm_mhz") or 1900.0)
        peak_tops = 2 * 8192 * 148 * sm_mhz * 1e6 * 1e-12
        int8_pipe = {"executed_tops": 2 * macs / (kern_mean_ms * 1e-3) * 1e-12, "peak_tops": peak_tops,
                     "frac": 2 * macs / (kern_mean_ms * 1e-3) * 1e-12 / peak_tops, "ctas": ctas, "sms": 148,
                     "slices": 5, "slice_pairs": 15,
                     "note": "int8 tensor-pipe utilisation of the emulation: a CTA's stage is MMA phase (~45 %) then drain + "
                             "RK4 combine + re-slicing of the next stage vector (serial per column tile), and 4096 columns "
                             "fill 128 of the 148 SMs; peak = measured 8192 MAC/clk/SM x 148 SMs x SM clock under load"}
        traffic = None
        tprof = os.path.join(ROOT, "profiles", "rk4_ozaki_traffic.json")
        if os.path.exists(tprof):
            try:
                pj = json.load(open(tprof))
                traffic = float(pj["dram_bytes_fixed"]) + float(pj["dram_bytes_per_table_entry"]) * (2 * S_k + 1)
            except Exception:  # noqa: BLE001
                traffic = None

    # parity of the TIMED configuration: final states of the e2e solve (1000 steps) on 32 chosen columns against the
    # answers of the UNMODIFIED reference on the same inputs (tests/golden/fullsize.npz, generated by make_golden.py)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "fullsize.npz"))
    cols = fx["cfg4_cols"]
    got = case["out_host"].numpy()[:, cols]
    parity = {"max_col_l2": float(np.max(np.linalg.norm(got - fx["cfg4_y"], axis=0))), "columns": int(cols.size), "rk4_steps": S,
              "against": "tests/golden/fullsize.npz: final states of solve_lmde(method='RK4') of the unmodified reference on "
                         "these 32 columns of the same batch (first / last octet, cluster-shared octet, random)",
              "bar": 1e-8, "max_unitarity_drift": float(np.max(np.abs(np.linalg.norm(case["out_host"].numpy(), axis=0) - 1.0)))}

    # ---------------- cfg3: vectorised Lindblad 729, batch 4096, scipy_expm, T = 0.2, max_dt = 1e-2 ----------------
    def cfg3_record():
        H0c, Hsc, Lsc, Yc, sigc = W.cfg3()
        m = Hsc.shape[-1] ** 2
        model3 = qd.LindbladModel(static_hamiltonian=H0c, hamiltonian_operators=Hsc,
                                  hamiltonian_signals=[qd.Signal(a, nu, ph) for a, nu, ph in sigc], static_dissipators=Lsc,
                                  rotating_frame=np.diag(H0c).real, vectorized=True)
        yd = qd.asarray(Yc)
        steps3, kw3 = 20, dict(method="scipy_expm", max_dt=1e-2)
        best, res = float("inf"), None
        for it in range(4):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            res = qd.solve_lmde(model3, t_span=[0.0, 0.2], y0=yd, **kw3)
            a1.record()
            torch.cuda.synchronize()
            if it >= 1:
                best = min(best, a0.elapsed_time(a1))
        from qiskit_dynamics_b200.solvers.fixed_step import expm_squarings
        tmid = (np.arange(steps3) + 0.5) * 1e-2
        sq = expm_squarings(model3, model3._signal_table(tmid), 1e-2, 1)
        Kc = model3._collection().num_operators
        # algorithmic flops per step: generator formation + Pade-13-equivalent exponential + application
        f_gen = (4 * Kc + 6) * m * m
        f_expm = float(np.mean((6 + sq + 8.0 / 3) * 8 * m**3))
        f_apply = 8.0 * m * m * Yc.shape[1]
        alg = steps3 * (f_gen + f_expm + f_apply)
        cols3 = fx["cfg3_cols"]
        err3 = float(np.max(np.linalg.norm(res.y[-1][:, torch.from_numpy(cols3).to(dev)].cpu().numpy() - fx["cfg3_y"], axis=0)))
        return {"n": int(m), "K": int(Kc), "batch": int(Yc.shape[1]), "expm_steps": steps3, "ms": best, "ms_per_step": best / steps3,
                "state_rhs_per_s": Yc.shape[1] * steps3 / (best * 1e-3), "alg_tflops": alg / best * 1e-9,
                "alg_frac": alg / best * 1e-9 / peak_tf, "squarings_mean": float(np.mean(sq)),
                "flops_per_step": {"generator": f_gen, "expm": f_expm, "apply": f_apply},
                "parity_max_col_l2": err3, "parity_columns": int(cols3.size), "parity_against": "tests/golden/fullsize.npz (unmodified reference)",
                "gemm": "zgemm_ozaki_kernel<6> (int8 tensor cores, six byte slices per operand; QDB_ZGEMM_INT8=0: zgemm3m_kernel, fp64 DMMA)"
                        if os.environ.get("QDB_ZGEMM_INT8", "1") != "0" else "zgemm3m_kernel (fp64 DMMA)",
                "note": "solve_lmde(LindbladModel(vectorized=True), method='scipy_expm') from device-resident y0, best of 3 "
                        "after 1 warm-up, CUDA events, L2 flushed; one 'state RHS' here = one propagator application per column; "
                        "alg_frac = fp64-equivalent algorithmic flops over the fp64 DMMA roof (> 1: the products run as exact "
                        "int8 slice products on tcgen05)"}

    cfg3 = cfg3_record()

    # ---------------- sweep mode beside the headline (the shared-signal shortcut does not exist there) ----------------
    def sweep_rate(n2, K2, B2, S2, seed2):
        """Device-resident per-column-signal RK4 (cfg2 / cfg5-like): state-RHS/s and algorithmic fraction of the roof."""
        H0s, Hss, Ys, sigs = workload(n2, K2, 1, seed2)
        m2 = qd.HamiltonianModel(static_operator=H0s, operators=Hss, signals=[qd.Signal(a_, nu_, ph_) for a_, nu_, ph_ in sigs],
                                 rotating_frame=H0s)
        c2 = m2._collection()
        p_ops, p_stat = c2.packed()
        t2 = stage_time_grid(0.0, MAX_DT, S2)
        base = torch.from_numpy(m2._signal_table(t2)).to(dev)
        amp = 0.5 + torch.arange(B2, dtype=torch.float64, device=dev) / B2
        coeff2 = (base[:, :, None] * amp[None, None, :]).contiguous()  # (T, K, B): amplitude sweep
        y2_0 = m2.rotating_frame.state_into_frame_basis(qd.asarray(np.repeat(Ys, B2, axis=1)))
        y2 = y2_0.clone()

        def run():
            y2.copy_(y2_0)
            abi.rk4_steps(n2, c2.operators, c2.static_operator, p_ops, p_stat, coeff2, m2._frame_freqs(), t2, MAX_DT, y2, S2,
                          per_col=True)
        best = float("inf")
        for it in range(5):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            run()
            a1.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = min(best, a0.elapsed_time(a1))
        alg = S2 * B2 * (4 * ((4 * K2 + 8) * n2 * n2 + 12 * n2) + 28 * n2)
        tl = abi.rk4_tiling(n2, B2, K2)
        return {"n": n2, "K": K2, "batch": B2, "rk4_steps": S2, "us_per_rk4_step": best * 1e3 / S2,
                "state_rhs_per_s": 4.0 * S2 * B2 / (best * 1e-3), "alg_tflops": alg / best * 1e-9,
                "alg_frac": alg / best * 1e-9 / peak_tf, "tiling": tl,
                "kernel": "rk4_sweepf_kernel" if tl["m3"] == 2 else "rk4_sweep_kernel"}

    sweep_mode = {"cfg2": sweep_rate(32, 8, 1024, 1000, 2002), "cfg5_like": sweep_rate(81, 8, 8192, 64, 2005),
                  "note": "per-column signal values (parameter sweeps), device-resident, best of 3 after 2 warm-ups; "
                          "algorithmic flops per RHS = (4K+8) n^2 + 12 n (SURVEY 8(d))"}

    # ---------------- the shared-signal shortcut, stated beside the result (SURVEY 8(d) honesty note) ----------------
    y_dev = qd.asarray(case["Y"])
    sbest = float("inf")
    for it in range(3):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        qd.solve_lmde(case["model"], t_span=t_span, y0=y_dev, method="jax_RK4_parallel", max_dt=MAX_DT)
        a1.record()
        torch.cuda.synchronize()
        if it >= 1:
            sbest = min(sbest, a0.elapsed_time(a1))
    shortcut = {"method": "jax_RK4_parallel", "ms_per_step": sbest,
                "note": f"same {S} RK4 steps on the same batch via step propagators (shared signals only); reported for "
                        "transparency, not comparable with `value`, which counts direct per-column RHS evaluations"}

    # ---------------- CPU baseline: the reference itself on this box's host cores (bounded sample) ----------------
    threads = cpu_threads()
    cpu_rate, cpu_sec, kind = cpu_reference_rate(n, K, B, steps=2, warmup=1, rk4_steps=CPU_BASELINE_RK4_STEPS, threads=threads)
    cpu_rate1, _, _ = cpu_reference_rate(n, K, B, steps=1, warmup=0, rk4_steps=4, threads=1)

    line = {
        "metric": "rhs_evals_per_sec", "value": value, "unit": "state-RHS/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": CONFIG,
        "timing": {"l2": "flushed between timed iterations (256 MB write)", "rk4_steps_per_step": S,
                   "rhs_evals_per_column_per_step": 4 * S, "timed_region_s": (dev_ms + e2e_ms) * args.steps * 1e-3},
        "parity": parity,
        "e2e": {"value": e2e_value, "unit": "state-RHS/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(n * B * 16 + case["times"].size * 8),
                "d2h_bytes_per_step": int(n * B * 16 + 8 * B * world),
                "note": "solve_lmde(model, y0=pinned host) -> final state to pinned host + per-column squared norms "
                        "(outcome_prob_kernel; all_gather_into_tensor over NCCL when N > 1) to pinned host"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": kernel_name, "tiling": tiling, "achieved": achieved_tf, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                     "flops_per_launch": flops_launch, "kernel_ms": kern_mean_ms, "rk4_steps_per_launch": S_k,
                     "kernel_share_of_step": kern_mean_ms * (S / S_k) / dev_ms,
                     "peak_cublas_zgemm": cublas_tf,
                     "int8_pipe": int8_pipe, "fp64_kernel": fp64_kernel,
                     "note": ("achieved = ALGORITHMIC fp64 flops (8 per complex multiply-add, SURVEY 8(d)) / kernel time, "
                              "peak = the fp64 DMMA tensor roof.  The kernel leaves the fp64 pipe -- the contraction runs as "
                              "exact int8 slice products on tcgen05 (int8_pipe: executed int8 ops against the int8 roof) -- "
                              "so frac > 1 is the speed-up over a perfect fp64 tensor kernel, not a utilisation; "
                              "fp64_kernel is the DMMA stepper on the same table in the same run")
                     if int8_path else fp64_kernel["note"],
                     "peak_source": "live DMMA m8n8k4 issue-rate probe (qdb_dmma_probe); second witness peak_cublas_zgemm = "
                                    "torch.matmul complex128 4096^3 in the same run; MEASURED_PEAKS.json has no fp64 entry; "
                                    "B200 datasheet fp64 tensor 37-40 TFLOP/s"},
        "strong_scaling": strong,
        "cfg5": cfg5,
        "cfg3": cfg3,
        "sweep_mode": sweep_mode,
        "shared_signal_shortcut": shortcut,
        "cpu_baseline": {"value": cpu_rate, "unit": "state-RHS/s", "cores": threads, "kind": kind,
                         "sample": f"{CPU_BASELINE_RK4_STEPS} of the {S} RK4 steps of the same n={n}, K={K}, B={B} batch x 2 repeats "
                                   f"after 1 warm-up ({cpu_sec:.2f} s each), solve_lmde(method='RK4') of "
                                   + ("the unmodified reference (baseline/_ref)" if kind == "reference" else "the NumPy port (oracle/)")
                                   + f", {threads} BLAS threads",
                         "single_thread_value": cpu_rate1, "cpu_model": cpu_model(), "os_cpu_count": os.cpu_count(),
                         "blas": blas_info(), "numpy": np.__version__},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
