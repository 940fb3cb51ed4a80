"""Synthetic inputs of the five BASELINE.json configurations, as plain NumPy.

Shared by bench.py (both arms), the full-size parity tests (tests/test_fullsize_gpu.py) and the fixture
generator (tests/golden/make_golden.py), so that every one of them draws the same numbers.  Imports
nothing from the package or from oracle/.

    herm(n) = (A + A^dag) / (2 sqrt n),  A = N(0,1) + i N(0,1);  state columns complex normal, unit L2 norm;
    signals Signal(a_j, nu_j, phi_j) with a_j = 0.1 (j+1), nu_j = 0.2 j + 0.05, phi_j = 0.3 j.
"""
from __future__ import annotations

import numpy as np

MAX_DT = 1e-3


def _herm(rng, n):
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    return (a + a.conj().T) / (2 * np.sqrt(n))


def signal_params(K):
    return [(0.1 * (j + 1), 0.2 * j + 0.05, 0.3 * j) for j in range(K)]


def schrodinger(n, K, B, seed):
    """H0 = 5 herm(n), H_j = herm(n), B unit-norm columns, K (amp, freq, phase) triples."""
    rng = np.random.default_rng(seed)
    H0 = 5 * _herm(rng, n)
    Hs = np.array([_herm(rng, n) for _ in range(K)])
    Y = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    Y = Y / np.linalg.norm(Y, axis=0, keepdims=True)
    return H0, Hs, Y, signal_params(K)


def cfg4(B=4096, seed=2004):
    """Headline: 7-qubit dim-128 Schrodinger in the rotating frame of H0, K = 8, shared signals, RK4 max_dt = 1e-3."""
    return schrodinger(128, 8, B, seed)


def cfg2(B=1024, seed=2002):
    """5-qubit dim-32, 8 drive operators, amplitude sweep a_{b,j} = a_j (0.5 + b/B) on ONE replicated initial state.
    Returns (H0, Hs, y0 (n,), per-column list of K (amp, freq, phase) triples)."""
    H0, Hs, Y, sig = schrodinger(32, 8, 1, seed)
    cols = [[(a * (0.5 + b / B), nu, ph) for a, nu, ph in sig] for b in range(B)]
    return H0, Hs, Y[:, 0].copy(), cols


def cfg3(B=4096, seed=2003, n=27, K=3, n_diss=6):
    """3-transmon Lindblad dim 27 (vec-rho 729), K Hermitian drive operators, 6 static dissipators 0.05 N(0,1) real,
    B pure-state density matrices column-stacked; frame = diag(H0) (1-d); scipy_expm max_dt = 1e-2, T = 0.2."""
    rng = np.random.default_rng(seed)
    H0 = 5 * _herm(rng, n)
    Hs = np.array([_herm(rng, n) for _ in range(K)])
    Ls = 0.05 * rng.standard_normal((n_diss, n, n))
    psi = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    psi = psi / np.linalg.norm(psi, axis=0, keepdims=True)
    rho = np.einsum("ib,kb->ikb", psi, psi.conj())
    Y = rho.reshape(n * n, B, order="F")  # vec_F: row index i + k n
    return H0, Hs, Ls, Y, signal_params(K)


# ---------------------------------------------------------------------------------------------
# cfg5: DynamicsBackend-style pulse sweep (synthetic stand-in: qiskit is absent)
# ---------------------------------------------------------------------------------------------

CFG5_DT = 0.222  # sample width in ns (1 / 4.5 GHz)


def cfg5_system(levels=3, nq=4):
    """Chain of `nq` Duffing transmons with exchange coupling (docs/tutorials/dynamics_backend.rst:55-110 of the
    reference), 4 drive + 4 control channels.  Returns (H0, ops (8, n, n), carrier frequencies (8,))."""
    a = np.diag(np.sqrt(np.arange(1, levels)), 1).astype(complex)
    N = a.conj().T @ a
    eye = np.eye(levels, dtype=complex)

    def op(single, k):
        mats = [eye] * nq
        mats[nq - 1 - k] = single
        out = mats[0]
        for m in mats[1:]:
            out = np.kron(out, m)
        return out

    w = 2 * np.pi * np.array([5.0, 5.1, 4.9, 5.05])[:nq]
    alpha, J = 2 * np.pi * -0.33, 2 * np.pi * 0.002
    H0 = sum(w[k] * op(N, k) + 0.5 * alpha * op(N @ (N - eye), k) for k in range(nq))
    H0 = H0 + sum(J * (op(a, k) @ op(a.conj().T, k + 1) + op(a.conj().T, k) @ op(a, k + 1)) for k in range(nq - 1))
    drives = [2 * np.pi * 0.02 * op(a + a.conj().T, k) for k in range(nq)]
    ops = np.array(drives + drives)
    freqs = np.array(list(w / (2 * np.pi)) + list(np.roll(w, 1) / (2 * np.pi)))
    return H0, ops, freqs


def cfg5_point(k, nsim, nsamp, dt=CFG5_DT):
    """Sweep point k of nsim: Gaussian-square envelopes with swept amplitude and width, one (samples, phase) per channel.
    Returns a list of 8 (samples (nsamp,) complex, phase) pairs."""
    t = (np.arange(nsamp) + 0.5) * dt
    amp = 0.2 + 0.8 * k / nsim
    width = (0.2 + 0.6 * ((7 * k) % nsim) / nsim) * nsamp * dt
    c, rise = t[-1] / 2 + dt / 2, 0.15 * nsamp * dt
    out = []
    for j in range(8):
        env = amp * (1 + 0.1 * j) * np.exp(-0.5 * (np.clip((np.abs(t - c) - width / 2) / rise, 0.0, None)) ** 2)
        out.append((env.astype(complex), 0.1 * j))
    return out


def cfg5_measurement(levels=3, nq=4):
    """Every transmon measured into its own memory slot: (subsystem_dims, measurement_subsystems, memory_slot_indices)."""
    return [levels] * nq, list(range(nq)), list(range(nq))


# ---------------------------------------------------------------------------------------------
# columns on which the full-size GPU solves are compared with the oracle / the reference fixtures
# ---------------------------------------------------------------------------------------------


def parity_columns(B, count=32, seed=99):
    """First and last octet, the octet a 2-CTA cluster shares at the headline tiling (columns 24..31), the ragged
    tail, and random columns -- `count` distinct sorted indices."""
    want = list(range(min(8, B))) + list(range(max(0, B - 8), B)) + [c for c in range(24, 32) if c < B]
    picked = sorted(set(want))
    rng = np.random.default_rng(seed)
    while len(picked) < min(count, B):
        c = int(rng.integers(0, B))
        if c not in picked:
            picked.append(c)
    return np.array(sorted(picked[:max(count, len(set(want)))]), dtype=np.int64)
