"""Where does the host time of a cfg5 sweep through the public API go?  cProfile of Solver.solve on 8192 lists of 8
DiscreteSignals (n = 81, 64 RK4 steps) + FinalStateMeasurement, after a warm-up call."""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_workloads as W
import qiskit_dynamics_b200 as qd

nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nsamp = 64
H0, ops, freqs = W.cfg5_system()
solver = qd.Solver(static_hamiltonian=H0, hamiltonian_operators=list(ops), rotating_frame=H0)
dims, msub, mslots = W.cfg5_measurement()
meas = qd.FinalStateMeasurement(solver.model, subsystem_dims=dims, measurement_subsystems=msub, memory_slot_indices=mslots, max_outcome_level=1)
t0 = time.perf_counter()
lists = [[qd.DiscreteSignal(dt=W.CFG5_DT, samples=s, carrier_freq=float(freqs[j]), phase=ph) for j, (s, ph) in enumerate(W.cfg5_point(k, nsim, nsamp))] for k in range(nsim)]
print("build signal objects: %.3f s" % (time.perf_counter() - t0))
y0 = np.zeros(81, dtype=complex); y0[0] = 1.0
tf = nsamp * W.CFG5_DT
def run():
    out = solver.solve(t_span=[0, tf], y0=y0, signals=lists, method="RK4", max_dt=W.CFG5_DT)
    finals = out.final_states
    P = meas.probabilities(tf, finals)
    torch.cuda.synchronize()
    return P
run()
for _ in range(2):
    t0 = time.perf_counter(); run(); print("call: %.4f s" % (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
