"""GPU parity tests of the C-ABI kernels (libqdb.so through ctypes) against the NumPy oracle.

Tolerance: the north-star bar is max-over-columns L2 error < 1e-8 on final states; true fp64
DMMA arithmetic lands at 1e-13..1e-15, so the tests assert 1e-11 (solves) / 1e-12 (single ops)
to leave room only for summation-order differences.
"""
import numpy as np
import pytest
import scipy.linalg

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import numpy_oracle as orc  # noqa: E402
from conftest import max_col_l2  # noqa: E402

TOL_OP = 1e-12
TOL_SOLVE = 1e-11


@pytest.fixture(scope="module")
def abi():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from qiskit_dynamics_b200 import _abi
    _abi.lib()
    return _abi


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def model_inputs(n, K, B, seed, frame="full"):
    H0, Hs, Y, sig = orc.synthetic_schrodinger(n, K, B, seed)
    fr = {"full": H0, "diag": np.diag(H0).real, "none": None}[frame]
    Gd, G, d, U = orc.generator_model_operators(H0, Hs, fr)
    yfb = Y if U is None else U.conj().T @ Y
    specs = [orc.SigSpec(a, nu, ph) for (a, nu, ph) in sig]
    mu = None if d is None else -np.imag(d)
    return Gd, G, d, mu, yfb, specs


def test_pack_and_generator(abi):
    rng = np.random.default_rng(0)
    for n, K, T in ((5, 2, 3), (8, 1, 1), (27, 3, 4), (128, 8, 5)):
        ops = rng.standard_normal((K, n, n)) + 1j * rng.standard_normal((K, n, n))
        stat = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        coeff = rng.standard_normal((T, K))
        mu = rng.standard_normal(n) * 3
        times = rng.uniform(0, 5, T)
        d = -1j * mu
        ref = np.array([orc.operator_into_frame(d, t, orc.collection_evaluate(c, ops, stat)) for t, c in zip(times, coeff)])
        out = abi.generator(n, dev(ops), dev(stat), dev(coeff), dev(mu), dev(times), scale=0.5)
        np.testing.assert_allclose(out.cpu().numpy().reshape(T, n, n), 0.5 * ref, rtol=0, atol=TOL_OP)
        # packed layout round trip
        npad, kpad = abi.npad(n), (n + 15) // 16 * 16
        pk_ops, pk_stat = abi.pack_operators(dev(ops)), abi.pack_operators(dev(stat[None]))[0]
        outp = abi.generator(n, pk_ops, pk_stat, dev(coeff), dev(mu), dev(times), layout=abi.LAYOUT_PACKED).cpu().numpy()
        r, c = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        idx = ((r // 8) * (kpad // 4) + c // 4) * 32 + (r % 8) * 4 + c % 4
        np.testing.assert_allclose(outp[:, idx], ref, rtol=0, atol=TOL_OP)
        mask = np.ones(npad * kpad, bool)
        mask[idx.ravel()] = False
        assert np.all(outp[:, mask] == 0)
        # complex coefficients, no frame, no static (reference test_operator_collections.py:82-94)
        cc = rng.standard_normal((1, K)) + 1j * rng.standard_normal((1, K))
        outc = abi.generator(n, dev(ops), None, dev(cc), None, None).cpu().numpy().reshape(n, n)
        np.testing.assert_allclose(outc, np.tensordot(cc[0], ops, axes=1), rtol=0, atol=TOL_OP)
        # static only
        outs = abi.generator(n, None, dev(stat), None, None, None).cpu().numpy().reshape(n, n)
        np.testing.assert_allclose(outs, stat, rtol=0, atol=0)


@pytest.mark.parametrize("M,N,Kd", [(1, 1, 1), (5, 3, 7), (64, 64, 16), (65, 67, 17), (128, 4096, 128), (729, 40, 729), (200, 130, 33),
                                    # 3-product kernel (>= 74 tiles of 64 x 64, k >= 64): full tiles, ragged edges in all three dimensions
                                    (729, 729, 729), (200, 1500, 67), (321, 963, 130), (64, 4800, 64)])
def test_zgemm(abi, M, N, Kd):
    rng = np.random.default_rng(M * 1000 + N)
    A = rng.standard_normal((M, Kd)) + 1j * rng.standard_normal((M, Kd))
    Bm = rng.standard_normal((Kd, N)) + 1j * rng.standard_normal((Kd, N))
    C0 = rng.standard_normal((M, N)) + 1j * rng.standard_normal((M, N))
    out = abi.zgemm(dev(A), dev(Bm)).cpu().numpy()
    scale = np.sqrt(Kd)
    np.testing.assert_allclose(out, A @ Bm, rtol=0, atol=TOL_OP * scale * 10)
    alpha, beta = 0.3 - 1.2j, -0.7 + 0.4j
    cs = rng.standard_normal(N)
    pre = np.exp(1j * rng.standard_normal(Kd))
    post = np.exp(1j * rng.standard_normal(M))
    c = dev(C0)
    abi.zgemm(dev(A), dev(Bm), out=c, alpha=alpha, beta=beta, colscale=dev(cs), pre=dev(pre), post=dev(post))
    ref = beta * C0 + alpha * cs[None, :] * post[:, None] * (A @ (pre[:, None] * Bm))
    np.testing.assert_allclose(c.cpu().numpy(), ref, rtol=0, atol=TOL_OP * scale * 10)


def test_zgemm_three_product_vs_four_product(abi, monkeypatch):
    """The 3-product kernel changes rounding only: against the 4-product kernel (QDB_ZGEMM_4M=1) on the shapes of the
    vectorised-Lindblad propagator (729^3) and its application (729 x 4096 x 729) the difference stays at a few
    ulp of |A||B| (normwise bound), and both agree with NumPy.  (QDB_ZGEMM_INT8=0: the fp64 DMMA kernels are the subject.)"""
    monkeypatch.setenv("QDB_ZGEMM_INT8", "0")
    rng = np.random.default_rng(7)
    for M, N, Kd in ((729, 729, 729), (729, 4096, 729), (128, 4096, 128)):
        A = rng.standard_normal((M, Kd)) + 1j * rng.standard_normal((M, Kd))
        Bm = rng.standard_normal((Kd, N)) + 1j * rng.standard_normal((Kd, N))
        Ad, Bd = dev(A), dev(Bm)
        before = abi.launch_count()
        out3 = abi.zgemm(Ad, Bd).cpu().numpy()
        monkeypatch.setenv("QDB_ZGEMM_4M", "1")
        out4 = abi.zgemm(Ad, Bd).cpu().numpy()
        monkeypatch.delenv("QDB_ZGEMM_4M")
        assert 2 <= abi.launch_count() - before <= 3  # 3-product kernel (+ its split-k tail launch) and 4-product kernel
        bound = np.abs(A) @ np.abs(Bm)
        assert np.max(np.abs(out3 - out4) / bound) < 16 * np.finfo(float).eps
        assert np.max(np.abs(out3 - out4)) > 0  # two different kernels did run
        np.testing.assert_allclose(out3, A @ Bm, rtol=0, atol=TOL_OP * np.sqrt(Kd) * 10)


@pytest.mark.parametrize("M,N,Kd", [(729, 4096, 729), (264, 4096, 264), (729, 1000, 300), (300, 9600, 257)])
def test_zgemm_split_k_tail(abi, monkeypatch, M, N, Kd):
    """Products whose 64 x 64 tile count leaves the last wave less than half full run that wave as clusters of 2 / 4 / 8
    CTAs splitting k (deterministic DSMEM reduction): same result as the one-CTA-per-tile launch up to the rounding of a
    different summation order, with every epilogue option, and bit-identical from run to run.  (QDB_ZGEMM_INT8=0: the fp64
    DMMA kernels are the subject.)"""
    monkeypatch.setenv("QDB_ZGEMM_INT8", "0")
    rng = np.random.default_rng(M + N + Kd)
    A = dev(rng.standard_normal((M, Kd)) + 1j * rng.standard_normal((M, Kd)))
    Bm = dev(rng.standard_normal((Kd, N)) + 1j * rng.standard_normal((Kd, N)))
    C0 = dev(rng.standard_normal((M, N)) + 1j * rng.standard_normal((M, N)))
    kw = dict(alpha=0.3 - 1.2j, beta=-0.7 + 0.4j, colscale=dev(rng.standard_normal(N)), pre=dev(np.exp(1j * rng.standard_normal(Kd))),
              post=dev(np.exp(1j * rng.standard_normal(M))))
    before = abi.launch_count()
    c1 = abi.zgemm(A, Bm, out=C0.clone(), **kw)
    launches = abi.launch_count() - before
    c1b = abi.zgemm(A, Bm, out=C0.clone(), **kw)
    monkeypatch.setenv("QDB_ZGEMM_NO_TAIL", "1")
    before = abi.launch_count()
    c2 = abi.zgemm(A, Bm, out=C0.clone(), **kw)
    assert abi.launch_count() - before == 1
    monkeypatch.delenv("QDB_ZGEMM_NO_TAIL")
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    tail = tiles % sms
    assert launches == (2 if (tiles > sms and 0 < tail <= sms // 2) else 1)
    assert torch.equal(c1, c1b)
    bound = (A.abs() @ Bm.abs()).max().item()
    assert (c1 - c2).abs().max().item() < 32 * np.finfo(float).eps * bound
    ref = kw["beta"] * C0 + kw["alpha"] * kw["colscale"][None, :] * kw["post"][:, None] * (A @ (kw["pre"][:, None] * Bm))
    assert (c1 - ref).abs().max().item() < TOL_OP * np.sqrt(Kd) * 10


@pytest.mark.parametrize("n,K,B,frame", [(8, 3, 5, "full"), (5, 2, 3, "diag"), (5, 2, 1, "none"), (128, 8, 64, "full"), (32, 8, 40, "full")])
def test_rhs_shared_and_sweep(abi, n, K, B, frame):
    Gd, G, d, mu, y, specs = model_inputs(n, K, B, 1234 + n, frame)
    mu_d = None if mu is None else dev(mu)
    for t in (0.0, 0.37, 2.5):
        c = orc.signal_list_values(specs, t)
        ref = orc.model_rhs(t, y, specs, G, Gd, d)
        out = abi.rhs(n, dev(G), dev(Gd), dev(c), mu_d, t, dev(y)).cpu().numpy()
        assert max_col_l2(out, ref) < TOL_OP * 10
        # no static operator
        ref2 = orc.model_rhs(t, y, specs, G, None, d)
        out2 = abi.rhs(n, dev(G), None, dev(c), mu_d, t, dev(y)).cpu().numpy()
        assert max_col_l2(out2, ref2) < TOL_OP * 10
        # per-column coefficients: column b scaled amplitudes
        amp = 0.5 + np.arange(B) / B
        cb = c[:, None] * amp[None, :]  # (K, B)
        ref3 = np.stack([orc.model_rhs(t, y[:, b], None, None, orc.collection_evaluate(cb[:, b], G, Gd), d) for b in range(B)], axis=-1)
        out3 = abi.rhs(n, dev(G), dev(Gd), dev(cb), mu_d, t, dev(y), per_col=True).cpu().numpy()
        assert max_col_l2(out3, ref3) < TOL_OP * 10


def _oracle_rk4(Gd, G, d, specs, y, t0, h, S):
    t = t0
    for _ in range(S):
        y = orc.rk4_step(lambda tt, yy: orc.model_rhs(tt, yy, specs, G, Gd, d), t, y, h)
        t = t + h
    return y


@pytest.mark.parametrize("n,K,B,S,frame", [
    (128, 8, 72, 20, "full"),    # headline shape, ragged column count
    (128, 8, 4096, 2, "full"),   # headline batch
    (5, 2, 3, 50, "diag"),       # odd dimension -> padding
    (4, 1, 1, 30, "full"),       # cfg1-like single column
    (32, 8, 100, 10, "full"),    # cfg2 dimension
    (88, 3, 40, 5, "full"),      # 11 row tiles (uneven warp load)
    (200, 2, 24, 3, "none"),     # MR = 4
    (256, 1, 16, 2, "full"),
])
def test_rk4_fused_shared(abi, n, K, B, S, frame):
    Gd, G, d, mu, y, specs = model_inputs(n, K, B, 77 + n, frame)
    t0, h = 0.1, 1e-3 if n >= 32 else 0.01
    times = orc.stage_time_grid(t0, h, S)
    coeff = orc.signal_list_values(specs, times)
    ref = _oracle_rk4(Gd, G, d, specs, y, t0, h, S)
    Gdev, Gd_dev = dev(G), dev(Gd)
    yd = dev(y)
    abi.rk4_steps(n, Gdev, Gd_dev, abi.pack_operators(Gdev), abi.pack_operators(Gd_dev[None])[0], dev(coeff),
                  None if mu is None else dev(mu), times, h, yd, S)
    assert max_col_l2(yd.cpu().numpy(), ref) < TOL_SOLVE
    # chunked step loop (tiny workspace) gives the same answer
    yd2 = dev(y)
    ws = torch.empty(abi.workspace_bytes(abi.WS_RK4, n, K, B, 2), dtype=torch.uint8, device="cuda")
    abi.rk4_steps(n, Gdev, Gd_dev, abi.pack_operators(Gdev), abi.pack_operators(Gd_dev[None])[0], dev(coeff),
                  None if mu is None else dev(mu), times, h, yd2, S, workspace=ws)
    assert torch.equal(yd, yd2)


@pytest.mark.parametrize("n,B,S", [
    (128, 4096, 3),   # headline: 74 clusters x 7 octets
    (128, 4000, 2),   # ragged batch: last cluster partly empty
    (128, 3001, 2),   # ragged inside an octet
    (100, 2500, 2),   # 13 row tiles: rank 1 owns 5 of its 8 shared row tiles
    (72, 6000, 2),    # 9 row tiles: rank 1 owns a single shared row tile
    (128, 8192, 1),   # two waves of clusters
])
def test_rk4_split_clusters_bit_identical(abi, n, B, S):
    """Split mode (2-CTA clusters exchanging one column octet over DSMEM) performs the same DMMA
    sequence per element as whole-column CTAs, so the two tilings must agree bit for bit; the
    whole-column tiling is the one checked against the oracle in test_rk4_fused_shared."""
    import os
    rng = np.random.default_rng(n + B)
    table = dev((rng.standard_normal((2 * S + 1, n, n)) + 1j * rng.standard_normal((2 * S + 1, n, n))) * (3.0 / np.sqrt(n)))
    packed = abi.pack_operators(table)
    y0 = dev(rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B)))
    old = os.environ.pop("QDB_NO_SPLIT", None)
    try:
        tiling = abi.rk4_tiling(n, B, -1)  # the 4-product kernel a PACKED table runs on
        if n != 72:
            assert tiling["split"] == 1 and tiling["m3"] == 0, tiling
        y1 = y0.clone()
        abi.rk4_table_steps(n, packed, 1e-2, y1, S)
        os.environ["QDB_NO_SPLIT"] = "1"
        assert abi.rk4_tiling(n, B, -1)["split"] == 0
        y2 = y0.clone()
        abi.rk4_table_steps(n, packed, 1e-2, y2, S)
    finally:
        os.environ.pop("QDB_NO_SPLIT", None)
        if old is not None:
            os.environ["QDB_NO_SPLIT"] = old
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    # and against a plain fp64 torch restatement of RK4 on the same table (first 16 columns + the last 8)
    cols = torch.cat([torch.arange(16), torch.arange(B - 8, B)]).cuda()
    y = y0[:, cols]
    h = 1e-2
    for s in range(S):
        G0, G1, G2 = table[2 * s], table[2 * s + 1], table[2 * s + 2]
        k1 = G0 @ y
        k2 = G1 @ (y + 0.5 * h * k1)
        k3 = G1 @ (y + 0.5 * h * k2)
        k4 = G2 @ (y + h * k3)
        y = y + (1.0 / 6) * h * (k1 + 2 * k2 + 2 * k3 + k4)
    err = torch.linalg.vector_norm(y1[:, cols] - y, dim=0).max().item()
    assert err < TOL_SOLVE


@pytest.mark.parametrize("n,B,S", [
    (128, 4096, 3),   # headline: static-geometry split kernel
    (128, 4000, 2),   # ragged batch
    (121, 3001, 2),   # same tiling, padded rows
    (100, 2500, 2),   # dynamic geometry, split, rank 1 owns 5 of 8 shared row tiles
    (128, 512, 2),    # pure row split: one octet per 2-CTA cluster, deep fragment ring (static geometry)
    (100, 300, 2),    # pure row split, dynamic geometry, rank 1 owns 5 of 8 row tiles
    (128, 5, 3),      # a single partial octet
    (121, 300, 3),    # row-split kernel with padded rows (n < 128)
    (125, 520, 5),    # row-split kernel, 65 octets, odd number of steps
    (128, 64, 6),     # row-split kernel, 8 clusters
    (128, 1024, 2),   # whole-column CTAs, one column tile per warp
    (64, 4096, 2),    # one row tile per warp
    (72, 100, 2),     # 9 row tiles on 8 row warps
    (57, 300, 2),     # smallest dimension the 3-product kernel takes
])
def test_rk4_three_product_kernel(abi, n, B, S):
    """rk4_shared3m_kernel (3 real DMMAs per complex tile product, PACKED3M table) against the 4-product kernel
    and a plain fp64 torch restatement of RK4.  3M changes rounding, not the algorithm: tolerance 1e-12 on the
    column L2 error after S steps of a norm-10 generator."""
    rng = np.random.default_rng(3 * n + B)
    A = rng.standard_normal((2 * S + 1, n, n)) + 1j * rng.standard_normal((2 * S + 1, n, n))
    table = dev((A - A.conj().transpose(0, 2, 1)) * (5.0 / np.sqrt(2 * n)))
    packed = abi.pack_operators(table)
    packed3 = abi.to_packed3m(packed)
    y0 = dev(rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B)))
    h = 1e-2
    y3 = y0.clone()
    abi.rk4_table_steps(n, packed3, h, y3, S, layout=abi.LAYOUT_PACKED3M)
    y4 = y0.clone()
    abi.rk4_table_steps(n, packed, h, y4, S, layout=abi.LAYOUT_PACKED)
    scale = torch.linalg.vector_norm(y0, dim=0).max().item()
    assert torch.linalg.vector_norm(y3 - y4, dim=0).max().item() < 1e-12 * scale
    cols = torch.cat([torch.arange(min(16, B)), torch.arange(max(B - 8, 0), B)]).cuda()
    y = y0[:, cols]
    for s in range(S):
        G0, G1, G2 = table[2 * s], table[2 * s + 1], table[2 * s + 2]
        k1 = G0 @ y
        k2 = G1 @ (y + 0.5 * h * k1)
        k3 = G1 @ (y + 0.5 * h * k2)
        k4 = G2 @ (y + h * k3)
        y = y + (1.0 / 6) * h * (k1 + 2 * k2 + 2 * k3 + k4)
    assert torch.linalg.vector_norm(y3[:, cols] - y, dim=0).max().item() < 1e-12 * scale


@pytest.mark.parametrize("n,B,S", [(128, 512, 7), (128, 20, 3), (122, 333, 4)])
def test_rowsplit_kernel_against_previous_tiling(abi, n, B, S):
    """rk4_rowsplit3m_kernel (16-deep fragment ring, own k half first, mid-stage mbarrier wait) against
    rk4_shared3m_kernel<2,0,split> (QDB_ROWSPLIT_OLD=1): rank 0 of every cluster performs the identical DMMA sequence
    (rows 0..63 bit for bit), rank 1 sums the two k halves in the other order (rounding only)."""
    import os
    rng = np.random.default_rng(n + 7 * B)
    A = rng.standard_normal((2 * S + 1, n, n)) + 1j * rng.standard_normal((2 * S + 1, n, n))
    table = dev((A - A.conj().transpose(0, 2, 1)) * (5.0 / np.sqrt(2 * n)))
    packed3 = abi.to_packed3m(abi.pack_operators(table))
    y0 = dev(rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B)))
    assert abi.rk4_tiling(n, B)["col_tiles_per_warp"] == 0 and abi.rk4_tiling(n, B)["split"] == 1
    old = os.environ.pop("QDB_ROWSPLIT_OLD", None)
    try:
        y_new = y0.clone()
        abi.rk4_table_steps(n, packed3, 1e-2, y_new, S, layout=abi.LAYOUT_PACKED3M)
        os.environ["QDB_ROWSPLIT_OLD"] = "1"
        y_old = y0.clone()
        abi.rk4_table_steps(n, packed3, 1e-2, y_old, S, layout=abi.LAYOUT_PACKED3M)
    finally:
        os.environ.pop("QDB_ROWSPLIT_OLD", None)
        if old is not None:
            os.environ["QDB_ROWSPLIT_OLD"] = old
    torch.cuda.synchronize()
    scale = torch.linalg.vector_norm(y0, dim=0).max().item()
    assert torch.linalg.vector_norm(y_new - y_old, dim=0).max().item() < 1e-13 * scale
    if S == 1:
        assert torch.equal(y_new[:64], y_old[:64])


def test_generator_packed3m_layout(abi):
    rng = np.random.default_rng(11)
    n, K, T = 100, 3, 4
    ops = rng.standard_normal((K, n, n)) + 1j * rng.standard_normal((K, n, n))
    stat = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    coeff, mu, times = rng.standard_normal((T, K)), rng.standard_normal(n), rng.uniform(0, 3, T)
    pk_ops, pk_stat = abi.pack_operators(dev(ops)), abi.pack_operators(dev(stat[None]))[0]
    a = abi.generator(n, pk_ops, pk_stat, dev(coeff), dev(mu), dev(times), layout=abi.LAYOUT_PACKED)
    b = abi.generator(n, pk_ops, pk_stat, dev(coeff), dev(mu), dev(times), layout=abi.LAYOUT_PACKED3M)
    assert b.shape == (T, abi.packed_elems(n) * 3 // 2)
    assert torch.equal(b, abi.to_packed3m(a))


@pytest.mark.parametrize("n,K,B,S,frame", [
    (32, 8, 48, 10, "full"),    # cfg2 shape: small-operator kernel, 4 row warps x 2 operator groups
    (5, 2, 3, 20, "diag"),      # one row tile, partial octet
    (128, 8, 40, 3, "full"),    # generic kernel
    (16, 2, 600, 4, "none"),    # two row tiles, 75 CTAs
    (27, 3, 20, 6, "full"),     # padded rows inside 4 row tiles
    (17, 5, 9, 5, "diag"),      # 3 row tiles, k padded to 32
    (8, 1, 10, 8, "none"),      # single operator group
    (12, 16, 5, 3, "full"),     # most operators the small kernel takes
    (24, 17, 4, 2, "full"),     # one more: generic kernel
    (81, 8, 24, 3, "full"),     # cfg5 dimension: 21 data k-tiles of 24 (padded pass list)
    (50, 3, 16, 3, "diag"),     # 13 data k-tiles of 16
])
def test_rk4_fused_sweep(abi, n, K, B, S, frame):
    Gd, G, d, mu, y, specs = model_inputs(n, K, B, 99 + n, frame)
    t0, h = 0.0, 1e-3 if n >= 32 else 0.01
    times = orc.stage_time_grid(t0, h, S)
    base = orc.signal_list_values(specs, times)  # (T, K)
    amp = 0.5 + np.arange(B) / B
    coeff = base[:, :, None] * amp[None, None, :]  # (T, K, B)
    ref = np.empty_like(y)
    for b in range(B):
        sp = [orc.SigSpec(s.envelope * amp[b], s.carrier_freq, s.phase) for s in specs]
        ref[:, b] = _oracle_rk4(Gd, G, d, sp, y[:, b], t0, h, S)
    Gdev, Gd_dev = dev(G), dev(Gd)
    yd = dev(y)
    abi.rk4_steps(n, Gdev, Gd_dev, abi.pack_operators(Gdev), abi.pack_operators(Gd_dev[None])[0], dev(coeff),
                  None if mu is None else dev(mu), times, h, yd, S, per_col=True)
    assert max_col_l2(yd.cpu().numpy(), ref) < TOL_SOLVE
    # no static operator
    ref2 = np.empty_like(y)
    for b in range(min(B, 4)):
        sp = [orc.SigSpec(s.envelope * amp[b], s.carrier_freq, s.phase) for s in specs]
        ref2[:, b] = _oracle_rk4(None, G, d, sp, y[:, b], t0, h, S)
    yd = dev(y)
    abi.rk4_steps(n, Gdev, None, abi.pack_operators(Gdev), None, dev(coeff), None if mu is None else dev(mu), times, h, yd, S, per_col=True)
    assert max_col_l2(yd.cpu().numpy()[:, :4], ref2[:, :4]) < TOL_SOLVE


@pytest.mark.parametrize("kernel", ["formed", "legacy"])
@pytest.mark.parametrize("n,K,B,S,frame", [
    (32, 8, 48, 6, "full"),     # cfg2 shape: 4 row warps, two DMMA k-steps over the operator index
    (81, 8, 40, 3, "full"),     # cfg5 dimension: odd n (matrix columns padded to 82), 11 row tiles on 4 x 3 slots
    (128, 8, 72, 3, "full"),    # cfg4 shape in sweep mode: 8 row warps x 2 row tiles, ragged last CTA
    (5, 3, 3, 10, "diag"),      # one row tile, one k-step, partial octet
    (17, 5, 9, 5, "diag"),      # K padded 5 -> 8
    (50, 4, 16, 3, "none"),     # exactly one k-step, no frame
    (27, 6, 20, 4, "full"),
    (200, 3, 12, 2, "full"),    # 25 row tiles: 8 row warps x 4 row tiles
    (16, 7, 600, 3, "none"),    # many CTAs of a small system
    (12, 16, 5, 3, "full"),     # four DMMA k-steps over the operator index
    (24, 11, 7, 3, "diag"),     # three k-steps, K padded 11 -> 12
    (81, 12, 16, 2, "none"),    # three k-steps at the cfg5 dimension (largest register footprint)
    (9, 2, 20, 5, "diag"),      # K < 3: formed only when pinned
    (40, 1, 10, 4, "full"),
])
def test_rk4_sweep_kernel_variants(abi, monkeypatch, kernel, n, K, B, S, frame):
    """Both sweep-kernel families on the same inputs: the formed-generator kernel (DMMA over the operator index + DFMA
    product, rk4_sweepf.cu) and the operator-pass kernels (rk4_sweep_kernel / rk4_sweep_small_kernel), pinned through
    QDB_SWEEP_KERNEL, each against the oracle's per-column solves; with and without a static operator."""
    monkeypatch.setenv("QDB_SWEEP_KERNEL", kernel)
    assert abi.rk4_tiling(n, B, K)["m3"] == (2 if kernel == "formed" else 0)
    Gd, G, d, mu, y, specs = model_inputs(n, K, B, 77 + n, frame)
    t0, h = 0.0, 1e-3 if n >= 32 else 0.01
    times = orc.stage_time_grid(t0, h, S)
    base = orc.signal_list_values(specs, times)
    amp = 0.5 + np.arange(B) / B
    coeff = base[:, :, None] * amp[None, None, :]
    nref = min(B, 24)
    cols = np.unique(np.concatenate([np.arange(min(B, 12)), np.arange(B - min(B, 12), B)]))[:nref]
    Gdev, Gd_dev = dev(G), dev(Gd)
    for static in (True, False):
        ref = {}
        for b in cols:
            sp = [orc.SigSpec(s.envelope * amp[b], s.carrier_freq, s.phase) for s in specs]
            ref[b] = _oracle_rk4(Gd if static else None, G, d, sp, y[:, b], t0, h, S)
        yd = dev(y)
        abi.rk4_steps(n, Gdev, Gd_dev if static else None, abi.pack_operators(Gdev),
                      abi.pack_operators(Gd_dev[None])[0] if static else None, dev(coeff), None if mu is None else dev(mu),
                      times, h, yd, S, per_col=True)
        out = yd.cpu().numpy()
        assert max(np.linalg.norm(out[:, b] - ref[b]) for b in cols) < TOL_SOLVE
        assert np.all(np.isfinite(out))


def test_rk4_sweep_kernels_agree_full_batch(abi, monkeypatch):
    """cfg5-like batch (n=81, K=8, 2048 columns: 64 CTAs): every column of the formed-generator kernel against the
    operator-pass kernel -- two different summation orders of the same per-column solve."""
    n, K, B, S = 81, 8, 2048, 2
    Gd, G, d, mu, y, specs = model_inputs(n, K, B, 5, "full")
    times = orc.stage_time_grid(0.0, 1e-3, S)
    coeff = orc.signal_list_values(specs, times)[:, :, None] * (0.5 + np.arange(B) / B)[None, None, :]
    args = (n, dev(G), dev(Gd), abi.pack_operators(dev(G)), abi.pack_operators(dev(Gd)[None])[0], dev(coeff), dev(mu), times, 1e-3)
    outs = {}
    for kernel in ("formed", "legacy"):
        monkeypatch.setenv("QDB_SWEEP_KERNEL", kernel)
        yd = dev(y)
        abi.rk4_steps(*args, yd, S, per_col=True)
        outs[kernel] = yd.cpu().numpy()
    assert max_col_l2(outs["formed"], outs["legacy"]) < 1e-13
    assert np.max(np.abs(outs["formed"] - outs["legacy"])) > 0


@pytest.mark.parametrize("B,S", [(24, 3), (2048, 2)])  # 4-product GEMM stages / 3-product GEMM stages (160 tiles)
def test_rk4_generic_large_n(abi, B, S):
    n, K = 264, 2
    Gd, G, d, mu, y, specs = model_inputs(n, K, B, 5, "full")
    t0, h = 0.0, 1e-3
    times = orc.stage_time_grid(t0, h, S)
    coeff = orc.signal_list_values(specs, times)
    ref = _oracle_rk4(Gd, G, d, specs, y, t0, h, S)
    yd = dev(y)
    abi.rk4_steps(n, dev(G), dev(Gd), None, None, dev(coeff), dev(mu), times, h, yd, S)
    assert max_col_l2(yd.cpu().numpy(), ref) < TOL_SOLVE


@pytest.mark.parametrize("n,norm", [(2, 3.0), (9, 0.3), (27, 5.0), (64, 40.0), (200, 1.0)])
def test_expm_matches_scipy(abi, n, norm):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = A - A.conj().T  # anti-Hermitian-ish generator, keeps exp bounded
    A = A * (norm / np.linalg.norm(A, 1)) + 0.01 * rng.standard_normal((n, n))
    s = max(0, int(np.ceil(np.log2(np.linalg.norm(A, 1) / 0.7))))
    out = abi.expm(dev(A), s).cpu().numpy()
    ref = scipy.linalg.expm(A)
    assert np.max(np.abs(out - ref)) < 1e-12 * max(1.0, norm)


def test_expm_steps(abi):
    # vectorised Lindblad, n = 3 -> 9, with 1-d frame, against the oracle's scipy_expm stepping
    n, K, B = 3, 2, 4
    H0, Hs, Ls, Y, sig = orc.synthetic_lindblad(n, K, 4, B, 31)
    specs = [orc.SigSpec(a, nu, ph) for (a, nu, ph) in sig]
    Hd, Hops, Ds, Do, d, U = orc.lindblad_model_operators(H0, Hs, Ls, None, np.diag(H0).real)
    Sst, ops = orc.vectorized_lindblad_collection(Hd, Hops, Ds, Do)
    mu = orc.vec_frame_phase(d)
    t0, h, S = 0.0, 0.05, 10
    y = Y.copy()
    t = t0
    for _ in range(S):
        y = orc.expm_step(lambda tt: orc.vectorized_map_into_frame(d, tt, orc.collection_evaluate(orc.signal_list_values(specs, tt), ops, Sst)), t, y, h)
        t = t + h
    starts = orc.stage_time_grid(t0, h, S)[0::2][:-1]
    mids = starts + h / 2
    coeff = orc.signal_list_values(specs, mids)
    bound = abs(h) * (np.linalg.norm(Sst, 1) + np.abs(coeff) @ np.array([np.linalg.norm(o, 1) for o in ops]))
    sq = np.maximum(0, np.ceil(np.log2(np.maximum(bound, 1e-300) / 0.7))).astype(np.int32)
    yd = dev(Y)
    abi.expm_steps(n * n, dev(ops), dev(Sst), dev(coeff), dev(mu), mids, sq, h, yd, S)  # propagators of all steps batched
    assert max_col_l2(yd.cpu().numpy(), y) < TOL_SOLVE
    # the one-step-at-a-time route (workspace of the S = 1 size) and a chunked one (3 steps per chunk)
    m2 = n * n
    for steps_of_room in (1, 3):
        ws = torch.empty(abi.workspace_bytes(abi.WS_EXPM, m2, K, B, steps_of_room) + (S - steps_of_room) * 8 + 256,
                         dtype=torch.uint8, device="cuda")
        y2 = dev(Y)
        abi.expm_steps(m2, dev(ops), dev(Sst), dev(coeff), dev(mu), mids, sq, h, y2, S, workspace=ws)
        assert max_col_l2(y2.cpu().numpy(), y) < TOL_SOLVE
        assert max_col_l2(y2.cpu().numpy(), yd.cpu().numpy()) < 1e-13


def test_signal_table_device(abi):
    """Row f3: qdb_signal_table_f64 against the host SignalList (pinned to the reference by the golden signal
    fixtures).  Bin selection must be exact -- checked with samples that spell their own index, on stage-time
    grids that land on bin edges -- values to 1e-14."""
    import qiskit_dynamics_b200 as qd
    from qiskit_dynamics_b200.signals import compile_signal_program
    from qiskit_dynamics_b200.solvers import stage_time_grid
    rng = np.random.default_rng(5)
    for dt, h in ((0.1, 0.05), (1 / 4.5, 1 / 4.5), (0.25, 0.1), (0.2222222222222222, 0.1111111111111111)):
        N = 37
        times = np.concatenate([stage_time_grid(0.0, h, 90), [-0.3, -1e-18, N * dt, N * dt + 1.0, 0.7 - 3 * dt]])
        # (1) index exactness: envelope value == bin index + 1, no carrier
        ident = qd.SignalList([qd.DiscreteSignal(dt, np.arange(N) + 1.0), qd.DiscreteSignal(dt, np.arange(N) + 1.0, start_time=0.7)])
        got = compile_signal_program(ident).table(times, "cuda").cpu().numpy()
        assert np.array_equal(got, ident.table(times))
        # (2) carriers, phases, complex samples, constant terms, sums inside one channel
        sl = qd.SignalList([
            qd.DiscreteSignal(dt, rng.standard_normal(N) + 1j * rng.standard_normal(N), carrier_freq=1.3, phase=0.4),
            qd.DiscreteSignal(dt, rng.standard_normal(N), start_time=0.35, carrier_freq=-0.7) + qd.Signal(0.25, 2.0, -0.3) + 0.5,
            qd.Signal(0.8 - 0.2j, 0.05, 1.0),
            2.5,
        ])
        got = compile_signal_program(sl).table(times, "cuda").cpu().numpy()
        np.testing.assert_allclose(got, sl.table(times), rtol=0, atol=1e-14)
    # (3) sweep: per-column amplitudes (samples) and per-column carrier frequencies (parameters)
    B, N, dt = 70, 20, 0.1
    times = stage_time_grid(0.0, 0.05, 45)
    base = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    lists = [qd.SignalList([qd.DiscreteSignal(dt, base * (0.5 + b / B), carrier_freq=1.0 + 0.01 * b, phase=0.1),
                            qd.Signal(0.1 * (b + 1), 0.3, 0.0)]) for b in range(B)]
    prog = compile_signal_program(lists)
    assert prog.columns == B and prog.params_per_column and prog.samples_per_column
    got = prog.table(times, "cuda").cpu().numpy()
    want = np.stack([sl.table(times) for sl in lists], axis=-1)
    assert got.shape == (times.shape[0], 2, B)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-14)


def test_argument_errors(abi):
    y = torch.zeros((4, 2), dtype=torch.complex128, device="cuda")
    with pytest.raises(abi.QdbError):
        abi.rhs(4, None, None, None, None, 0.0, y)  # empty collection (operator_collections.py:119-122)
    with pytest.raises(abi.QdbError):
        abi.rhs(4, None, torch.zeros((4, 4), dtype=torch.complex128), None, None, 0.0, y)  # CPU tensor: no fallback
    with pytest.raises(abi.QdbError):
        abi.zgemm(torch.zeros((4, 3), dtype=torch.complex128, device="cuda"), torch.zeros((4, 3), dtype=torch.complex128, device="cuda"))
