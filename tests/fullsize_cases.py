"""The full-size BASELINE configurations as the parity tests use them: inputs from bench_workloads, the oracle's answer on
chosen columns (NumPy restatement of the reference), and the reference-generated fixture (tests/golden/fullsize.npz)."""
import numpy as np

import bench_workloads as W
from oracle import numpy_oracle as orc

CFG5_NSIM, CFG5_NSAMP = 8192, 64


def specs(sig):
    return [orc.SigSpec(a, nu, ph) for a, nu, ph in sig]


def oracle_cfg4(cols, t_end=1.0, B=4096):
    H0, Hs, Y, sig = W.cfg4(B)
    _, ys = orc.solve_hamiltonian(H0, Hs, specs(sig), H0, [0, t_end], Y[:, cols], W.MAX_DT)
    return ys[-1]


def oracle_cfg2(cols):
    H0, Hs, y0, per_col = W.cfg2()
    out = []
    for b in cols:
        _, ys = orc.solve_hamiltonian(H0, Hs, specs(per_col[int(b)]), H0, [0, 1.0], y0, W.MAX_DT)
        out.append(ys[-1])
    return np.stack(out, axis=-1)


def oracle_cfg3(cols):
    H0, Hs, Ls, Y, sig = W.cfg3()
    _, ys = orc.solve_vectorized_lindblad(H0, Hs, specs(sig), Ls, None, None, np.diag(H0).real, [0, 0.2], Y[:, cols], 1e-2)
    return ys[-1]


def cfg5_specs(k, freqs):
    return [orc.SigSpec(("discrete", W.CFG5_DT, smp, 0.0), float(freqs[j]), ph)
            for j, (smp, ph) in enumerate(W.cfg5_point(int(k), CFG5_NSIM, CFG5_NSAMP))]


def oracle_cfg5(cols):
    H0, ops, freqs = W.cfg5_system()
    y0 = np.zeros(H0.shape[0], dtype=complex)
    y0[0] = 1.0
    out = []
    for k in cols:
        _, ys = orc.solve_hamiltonian(H0, ops, cfg5_specs(k, freqs), H0, [0, CFG5_NSAMP * W.CFG5_DT], y0, W.CFG5_DT)
        out.append(ys[-1])
    return np.stack(out, axis=-1)
