"""zgemm_ozaki_kernel: the complex GEMM whose fp64 contraction is emulated on the int8 tensor cores (tcgen05.mma kind::i8;
six signed byte slices per operand against a per-row scale of A and a per-column scale of B, exact int32 slice products,
fp64 recombination).  Truncation 2^-48 of the row / column maximum per operand: tolerances below are normwise and stated
per assertion.  QDB_ZGEMM_INT8=2 forces the emulation on shapes the dispatcher would leave to the DMMA kernels."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from qiskit_dynamics_b200 import _abi
    _abi.lib()
    return _abi


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("M,N,K", [(729, 4096, 729), (729, 729, 729), (1, 1, 1), (5, 3, 7), (128, 32, 128), (129, 33, 129),
                                    (300, 1000, 200), (100, 37, 1000), (257, 4099, 385)])
def test_int8_gemm_matches_cublas_with_every_epilogue_option(abi, monkeypatch, M, N, K):
    monkeypatch.setenv("QDB_ZGEMM_INT8", "2")
    rng = np.random.default_rng(M + 3 * N + 7 * K)
    A = dev(rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K)))
    B = dev(rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N)))
    before = abi.launch_count()
    out = abi.zgemm(A, B)
    assert abi.launch_count() - before == 4  # two slicing passes over B, one over A, the product
    assert rel(out, A @ B) < 1e-12
    C0 = dev(rng.standard_normal((M, N)) + 1j * rng.standard_normal((M, N)))
    cs, pre, post = dev(rng.standard_normal(N)), dev(np.exp(1j * rng.standard_normal(K))), dev(np.exp(1j * rng.standard_normal(M)))
    alpha, beta = 0.3 - 1.2j, -0.7 + 0.4j
    c = abi.zgemm(A, B, out=C0.clone(), alpha=alpha, beta=beta, colscale=cs, pre=pre, post=post)
    ref = beta * C0 + alpha * cs[None, :] * post[:, None] * (A @ (pre[:, None] * B))
    assert rel(c, ref) < 1e-12
    # bit-reproducible
    c2 = abi.zgemm(A, B, out=C0.clone(), alpha=alpha, beta=beta, colscale=cs, pre=pre, post=post)
    assert torch.equal(c, c2)


def test_scales_are_per_row_of_a_and_per_column_of_b(abi, monkeypatch):
    """Rows of A from 1e-8 to 1e8 and columns of B from 1e-20 to 1e20: every element of C keeps a relative error of 1e-11
    against its own row x column scale (a uniformly scaled row or column loses nothing); zero rows / columns stay exact."""
    monkeypatch.setenv("QDB_ZGEMM_INT8", "2")
    M, N, K = 200, 96, 300
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K))) * np.logspace(-8, 8, M)[:, None]
    B = (rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))) * np.logspace(-20, 20, N)[None, :]
    A[17] = 0.0
    B[:, 5] = 0.0
    out = abi.zgemm(dev(A), dev(B)).cpu().numpy()
    ref = A @ B
    scale = np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=0)[None, :] * np.sqrt(K)
    scale[scale == 0] = 1.0
    assert np.max(np.abs(out - ref) / scale) < 1e-11
    assert np.all(out[17] == 0) and np.all(out[:, 5] == 0)


def test_five_slices_are_faster_and_coarser(abi, monkeypatch):
    monkeypatch.setenv("QDB_ZGEMM_INT8", "2")
    monkeypatch.setenv("QDB_ZGEMM_SLICES", "5")
    rng = np.random.default_rng(1)
    A = dev(rng.standard_normal((400, 500)) + 1j * rng.standard_normal((400, 500)))
    B = dev(rng.standard_normal((500, 640)) + 1j * rng.standard_normal((500, 640)))
    err = rel(abi.zgemm(A, B), A @ B)
    assert 1e-13 < err < 1e-9  # 2^-40 per operand


def test_dispatch_and_opt_out(abi, monkeypatch):
    """Default: products that fill half the SMs with 128 x 32 tiles and have k >= 384 take the emulation (4 launches), others
    the DMMA kernels (1 or 2); QDB_ZGEMM_INT8=0 keeps everything on the DMMA kernels."""
    rng = np.random.default_rng(2)
    A = dev(rng.standard_normal((729, 729)) + 1j * rng.standard_normal((729, 729)))
    B = dev(rng.standard_normal((729, 4096)) + 1j * rng.standard_normal((729, 4096)))
    monkeypatch.delenv("QDB_ZGEMM_INT8", raising=False)
    before = abi.launch_count()
    c_int8 = abi.zgemm(A, B)
    assert abi.launch_count() - before == 4
    before = abi.launch_count()
    abi.zgemm(A[:128, :128].contiguous(), B[:128].contiguous())
    assert abi.launch_count() - before <= 2
    monkeypatch.setenv("QDB_ZGEMM_INT8", "0")
    before = abi.launch_count()
    c_fp64 = abi.zgemm(A, B)
    assert abi.launch_count() - before <= 2
    assert rel(c_int8, c_fp64) < 1e-12


def test_generic_rk4_stage_takes_the_int8_gemm(abi, monkeypatch):
    """n > 256 shared-signal RK4 (the path of vectorised Lindblad models with RK4): every stage is one product with the RK4
    epilogue (k = G y_in, y_out = y + a k, acc += w k).  At n = 400, B = 2048 it runs on the int8 GEMM (per step: generator +
    4 x 4 launches + axpby); final states against the DMMA route of the same call < 1e-11 and against NumPy < 1e-11."""
    n, K, B, S, h = 400, 2, 2048, 3, 1e-3
    rng = np.random.default_rng(12)
    herm = lambda: (lambda a: (a + a.conj().T) / (2 * np.sqrt(n)))(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    G = np.array([-1j * herm() for _ in range(K)])
    Gd = -1j * 5 * herm()
    Y = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    Y /= np.linalg.norm(Y, axis=0, keepdims=True)
    times = np.arange(2 * S + 1) * (h / 2)
    coeff = np.stack([np.cos(3 * times), np.sin(2 * times)], axis=1)

    def run():
        y = dev(Y)
        before = abi.launch_count()
        abi.rk4_steps(n, dev(G), dev(Gd), None, None, dev(coeff), None, times, h, y, S)
        torch.cuda.synchronize()
        return y, abi.launch_count() - before

    monkeypatch.setenv("QDB_ZGEMM_INT8", "0")
    y_fp64, l_fp64 = run()
    monkeypatch.delenv("QDB_ZGEMM_INT8")
    y_int8, l_int8 = run()
    assert l_int8 > l_fp64  # four launches per product instead of one
    assert float(torch.linalg.vector_norm(y_int8 - y_fp64, dim=0).max()) < 1e-11
    y = Y[:, :8].copy()
    gen = lambda i: Gd + coeff[i, 0] * G[0] + coeff[i, 1] * G[1]
    for s in range(S):
        k1 = gen(2 * s) @ y
        k2 = gen(2 * s + 1) @ (y + 0.5 * h * k1)
        k3 = gen(2 * s + 1) @ (y + 0.5 * h * k2)
        k4 = gen(2 * s + 2) @ (y + h * k3)
        y = y + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    assert np.linalg.norm(y_int8.cpu().numpy()[:, :8] - y, axis=0).max() < 1e-11
