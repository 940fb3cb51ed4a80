"""Operator collections: the math objects  Lambda(c, y) = (G_d + sum_j c_j G_j) y.

Mirror of the reference's ``models/operator_collections.py`` protocol (attributes ``dim``,
``static_operator``, ``operators``; methods ``evaluate``, ``evaluate_rhs``, ``__call__``), backed
by HBM-resident complex128 tensors and the CUDA C-ABI.  This is the only collection type: the
``array_library`` selector of the reference is accepted for signature compatibility and must be
``None``/"torch"/"numpy" (dense); sparse and JAX variants do not exist here.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _abi
from ..arrays import asarray, asreal, CDTYPE, RDTYPE
from ..exceptions import QiskitError

_DENSE_LIBRARIES = (None, "torch", "numpy", "cuda")


def _check_library(array_library, who: str):
    if array_library not in _DENSE_LIBRARIES:
        raise QiskitError(
            f"{array_library} is not a valid array_library for {who}: the B200 build has a single dense "
            "CUDA implementation (no arraylias dispatch)."
        )


def _coefficients(c, device, K: int, what: str = "coefficients"):
    """-> (tensor (1,K) float64 or complex128)."""
    if c is None:
        return None
    if isinstance(c, torch.Tensor):
        t = c.to(device)
    else:
        t = torch.from_numpy(np.atleast_1d(np.asarray(c))).to(device)
    t = t.to(CDTYPE if t.is_complex() else RDTYPE).reshape(1, -1).contiguous()
    if t.shape[1] != K:
        raise QiskitError(f"{what} has length {t.shape[1]} but the collection holds {K} operators.")
    return t


def _as_columns(y: torch.Tensor):
    """(n,) -> (n,1); returns (y2d, restore)."""
    if y.ndim == 1:
        return y.reshape(-1, 1).contiguous(), (lambda out: out.reshape(-1))
    if y.ndim == 2:
        return y.contiguous(), (lambda out: out)
    raise QiskitError("state must be a vector (n,) or a column batch (n, B).")


class OperatorCollection:
    """(G_d + sum_j c_j G_j) y  (operator_collections.py:44-148)."""

    def __init__(self, static_operator=None, operators=None, array_library: Optional[str] = None):
        _check_library(array_library, "OperatorCollection")
        self._static_operator = asarray(static_operator)
        self._operators = asarray(operators)
        if self._operators is not None and self._operators.ndim == 2:
            self._operators = self._operators.unsqueeze(0).contiguous()
        self._packed = None
        self._norms = None

    @property
    def dim(self) -> int:
        if self._static_operator is not None:
            return int(self._static_operator.shape[-1])
        return int(self._operators.shape[-1])

    @property
    def static_operator(self):
        return self._static_operator

    @property
    def operators(self):
        return self._operators

    @property
    def num_operators(self) -> int:
        return 0 if self._operators is None else int(self._operators.shape[0])

    def _require_nonempty(self):
        if self._static_operator is None and self._operators is None:
            raise QiskitError(
                "OperatorCollection with None for both static_operator and operators cannot be evaluated."
            )

    # -- device layouts for the fused steppers ----------------------------------------------------
    def packed(self):
        """(ops_packed (K, npad^2) or None, static_packed (npad^2,) or None), built once."""
        if self._packed is None:
            ops_p = None if self._operators is None else _abi.pack_operators(self._operators)
            st_p = None if self._static_operator is None else _abi.pack_operators(self._static_operator.unsqueeze(0))[0]
            self._packed = (ops_p, st_p)
        return self._packed

    def norms1(self):
        """(||G_d||_1, [||G_j||_1]) as host floats: bound used to pick expm squarings without a sync
        inside the step loop."""
        if self._norms is None:
            s = 0.0 if self._static_operator is None else float(torch.linalg.matrix_norm(self._static_operator, 1))
            o = np.zeros(0) if self._operators is None else torch.linalg.matrix_norm(self._operators, 1).cpu().numpy()
            self._norms = (s, o)
        return self._norms

    # -- evaluation ---------------------------------------------------------------------------
    def evaluate(self, coefficients):
        """G_d + sum_j c_j G_j -> (n, n) tensor (operator_collections.py:101-122)."""
        self._require_nonempty()
        n = self.dim
        if self._operators is None:
            return self._static_operator
        c = _coefficients(coefficients, self._operators.device, self.num_operators)
        if c is None:
            raise QiskitError("coefficients required for a collection with operators.")
        return _abi.generator(n, self._operators, self._static_operator, c, None, None).reshape(n, n)

    def evaluate_rhs(self, coefficients, y):
        """(G_d + sum_j c_j G_j) y (operator_collections.py:124-134)."""
        self._require_nonempty()
        y2, restore = _as_columns(asarray(y))
        n = self.dim
        if y2.shape[0] != n:
            raise QiskitError(f"state has leading dimension {y2.shape[0]}, collection dimension is {n}.")
        c = None
        if self._operators is not None:
            c = _coefficients(coefficients, y2.device, self.num_operators)
            if c is None:
                raise QiskitError("coefficients required for a collection with operators.")
            if c.is_complex():  # complex coefficients: form G once, then one GEMM
                return restore(_abi.zgemm(self.evaluate(c), y2))
            c = c.reshape(-1)
        return restore(_abi.rhs(n, self._operators, self._static_operator, c, None, 0.0, y2))

    def __call__(self, coefficients, y=None):
        return self.evaluate(coefficients) if y is None else self.evaluate_rhs(coefficients, y)


# -------------------------------------------------------------------------------------------------
# Lindblad collections
# -------------------------------------------------------------------------------------------------


def vec_commutator(A: torch.Tensor) -> torch.Tensor:
    """-i (I kron A - A^T kron I), column stacking (models/model_utils.py:31-71); A (n,n) or (k,n,n)."""
    A = asarray(A)
    if A.ndim == 3:
        return torch.stack([vec_commutator(a) for a in A])
    iden = torch.eye(A.shape[-1], dtype=CDTYPE, device=A.device)
    return (-1j * (torch.kron(iden, A) - torch.kron(A.transpose(0, 1).contiguous(), iden))).contiguous()


def vec_dissipator(L: torch.Tensor) -> torch.Tensor:
    """conj(L) kron L - (I kron L^dag L + (L^dag L)^T kron I)/2 (models/model_utils.py:74-118)."""
    L = asarray(L)
    if L.ndim == 3:
        return torch.stack([vec_dissipator(x) for x in L])
    iden = torch.eye(L.shape[-1], dtype=CDTYPE, device=L.device)
    Lc = L.conj().resolve_conj()
    LdL = Lc.transpose(0, 1) @ L
    out = torch.kron(Lc, iden) @ torch.kron(iden, L) - 0.5 * (torch.kron(iden, LdL) + torch.kron(LdL.transpose(0, 1).contiguous(), iden))
    return out.contiguous()


class VectorizedLindbladCollection:
    """Lindblad generator as one (n^2, n^2) OperatorCollection (operator_collections.py:851-1061)."""

    def __init__(self, static_hamiltonian=None, hamiltonian_operators=None, static_dissipators=None,
                 dissipator_operators=None, array_library: Optional[str] = None):
        _check_library(array_library, "VectorizedLindbladCollection")
        self._static_hamiltonian = asarray(static_hamiltonian)
        self._hamiltonian_operators = asarray(hamiltonian_operators)
        self._static_dissipators = asarray(static_dissipators)
        self._dissipator_operators = asarray(dissipator_operators)

        static = None
        if self._static_hamiltonian is not None:
            static = vec_commutator(self._static_hamiltonian)
        if self._static_dissipators is not None:
            dsum = torch.sum(vec_dissipator(self._static_dissipators), dim=0)
            static = dsum if static is None else static + dsum
        parts = []
        if self._hamiltonian_operators is not None:
            parts.append(vec_commutator(self._hamiltonian_operators))
        if self._dissipator_operators is not None:
            parts.append(vec_dissipator(self._dissipator_operators))
        operators = torch.cat(parts, dim=0).contiguous() if parts else None
        self._operator_collection = OperatorCollection(static_operator=static, operators=operators)

    static_hamiltonian = property(lambda self: self._static_hamiltonian)
    hamiltonian_operators = property(lambda self: self._hamiltonian_operators)
    static_dissipators = property(lambda self: self._static_dissipators)
    dissipator_operators = property(lambda self: self._dissipator_operators)

    def evaluate_hamiltonian(self, ham_coefficients):
        return _evaluate_hamiltonian(self, ham_coefficients)

    def _concatenate_coefficients(self, ham_coefficients, dis_coefficients):
        has_h, has_d = self._hamiltonian_operators is not None, self._dissipator_operators is not None
        if has_h and has_d:
            return np.append(np.asarray(_host(ham_coefficients)), np.asarray(_host(dis_coefficients)), axis=-1)
        if has_h:
            return ham_coefficients
        if has_d:
            return dis_coefficients
        return None

    def evaluate(self, ham_coefficients, dis_coefficients):
        return self._operator_collection.evaluate(self._concatenate_coefficients(ham_coefficients, dis_coefficients))

    def evaluate_rhs(self, ham_coefficients, dis_coefficients, y):
        return self._operator_collection.evaluate_rhs(self._concatenate_coefficients(ham_coefficients, dis_coefficients), y)

    def __call__(self, ham_coefficients, dis_coefficients, y=None):
        if y is None:
            return self.evaluate(ham_coefficients, dis_coefficients)
        return self.evaluate_rhs(ham_coefficients, dis_coefficients, y)


def _host(c):
    return c.detach().cpu().numpy() if isinstance(c, torch.Tensor) else c


def _evaluate_hamiltonian(coll, ham_coefficients):
    """H_d + sum_j s_j H_j (operator_collections.py:406-432)."""
    if coll._static_hamiltonian is None and coll._hamiltonian_operators is None:
        raise QiskitError(
            f"{type(coll).__name__} with None for both static_hamiltonian and hamiltonian_operators cannot "
            "evaluate Hamiltonian."
        )
    return OperatorCollection(coll._static_hamiltonian, coll._hamiltonian_operators).evaluate(ham_coefficients)


class LindbladCollection:
    """Non-vectorised Lindblad RHS  (A+B) rho + rho (A-B) + sum_j g_j L_j rho L_j^dag
    with A = -1/2 sum (g_j) L_j^dag L_j, B = -iH (operator_collections.py:273-588).

    rho is (n, n) or a batch (l, n, n) (batch = LEADING axis, as in the reference).  Every
    product runs through the DMMA GEMM: the batch is folded into the GEMM's free dimension
    (right products as (l n, n) x (n, n); left products on the (n, l n) transposed view).
    """

    def __init__(self, static_hamiltonian=None, hamiltonian_operators=None, static_dissipators=None,
                 dissipator_operators=None, array_library: Optional[str] = None):
        _check_library(array_library, "LindbladCollection")
        self._static_hamiltonian = asarray(static_hamiltonian)
        self._hamiltonian_operators = asarray(hamiltonian_operators)
        self._static_dissipators = asarray(static_dissipators)
        self._dissipator_operators = asarray(dissipator_operators)
        self._ham = None
        if self._static_hamiltonian is not None or self._hamiltonian_operators is not None:
            self._ham = OperatorCollection(self._static_hamiltonian, self._hamiltonian_operators)
        self._static_product_sum = None
        self._static_adj = None
        if self._static_dissipators is not None:
            D = self._static_dissipators
            self._static_adj = D.conj().resolve_conj().transpose(-1, -2).contiguous()
            self._static_product_sum = (-0.5 * torch.sum(torch.matmul(self._static_adj, D), dim=0)).contiguous()
        self._dis_products = None
        self._dis_adj = None
        if self._dissipator_operators is not None:
            L = self._dissipator_operators
            self._dis_adj = L.conj().resolve_conj().transpose(-1, -2).contiguous()
            self._dis_products = OperatorCollection(self._static_product_sum,
                                                    (-0.5 * torch.matmul(self._dis_adj, L)).contiguous())

    static_hamiltonian = property(lambda self: self._static_hamiltonian)
    hamiltonian_operators = property(lambda self: self._hamiltonian_operators)
    static_dissipators = property(lambda self: self._static_dissipators)
    dissipator_operators = property(lambda self: self._dissipator_operators)

    def evaluate_hamiltonian(self, ham_coefficients):
        return _evaluate_hamiltonian(self, ham_coefficients)

    def evaluate(self, ham_coefficients, dis_coefficients):
        raise ValueError("Non-vectorized Lindblad collections cannot be evaluated without a state.")

    @staticmethod
    def _left(X: torch.Tensor, rho: torch.Tensor, out=None, beta=0.0) -> torch.Tensor:
        """X @ rho_b for every b: one GEMM on the (n, l*n) view."""
        l, n, _ = rho.shape
        cols = rho.permute(1, 0, 2).reshape(n, l * n).contiguous()
        res = _abi.zgemm(X, cols)
        return res.reshape(n, l, n).permute(1, 0, 2)

    @staticmethod
    def _right(rho: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
        """rho_b @ X for every b: one GEMM on the (l*n, n) view."""
        l, n, _ = rho.shape
        return _abi.zgemm(rho.reshape(l * n, n).contiguous(), X).reshape(l, n, n)

    # -- device operands of the fused kernels (csrc/lindblad.cu) --------------------------------------
    @property
    def dim(self) -> int:
        for ops in (self._static_hamiltonian, self._hamiltonian_operators, self._static_dissipators, self._dissipator_operators):
            if ops is not None:
                return int(ops.shape[-1])
        raise QiskitError("empty LindbladCollection")

    def fused_operands(self):
        """Operands of qdb_lindblad_*: rhs(X) = M1 X + X M2 + sum_j g_j L_j X L_j^dag with
        M1 = stat1 + sum_k c_k ops1[k], M2^T = stat2t + sum_k c_k ops2t[k], c = [ham coefficients, dis coefficients],
        ops1 = [-i H_k ; -1/2 L_k^dag L_k], ops2t = [(+i H_k)^T ; (-1/2 L_k^dag L_k)^T], and the dissipator stack
        [static D_j ; L_k] with g = [1, ..., dis coefficients].  Everything QDB_LAYOUT_PACKED, built once."""
        if getattr(self, "_fused", None) is None:
            n = self.dim
            dev = next(x for x in (self._static_hamiltonian, self._hamiltonian_operators, self._static_dissipators,
                                   self._dissipator_operators) if x is not None).device
            stat1 = torch.zeros((n, n), dtype=CDTYPE, device=dev)
            stat2 = torch.zeros((n, n), dtype=CDTYPE, device=dev)
            if self._static_hamiltonian is not None:
                stat1 = stat1 - 1j * self._static_hamiltonian
                stat2 = stat2 + 1j * self._static_hamiltonian
            if self._static_product_sum is not None:
                stat1 = stat1 + self._static_product_sum
                stat2 = stat2 + self._static_product_sum
            o1, o2 = [], []
            if self._hamiltonian_operators is not None:
                o1.append(-1j * self._hamiltonian_operators)
                o2.append(1j * self._hamiltonian_operators)
            if self._dissipator_operators is not None:
                prods = self._dis_products.operators
                o1.append(prods)
                o2.append(prods)
            ops1 = torch.cat(o1, dim=0).contiguous() if o1 else None
            ops2t = torch.cat(o2, dim=0).transpose(-1, -2).contiguous() if o2 else None
            diss = [x for x in (self._static_dissipators, self._dissipator_operators) if x is not None]
            diss = torch.cat(diss, dim=0).contiguous() if diss else None
            self._fused = dict(
                stat1=_abi.pack_operators(stat1.unsqueeze(0).contiguous())[0],
                stat2t=_abi.pack_operators(stat2.transpose(0, 1).contiguous().unsqueeze(0))[0],
                ops1=None if ops1 is None else _abi.pack_operators(ops1),
                ops2t=None if ops2t is None else _abi.pack_operators(ops2t),
                diss=None if diss is None else _abi.pack_operators(diss),
                n_static=0 if self._static_dissipators is None else int(self._static_dissipators.shape[0]),
                n_ham=0 if self._hamiltonian_operators is None else int(self._hamiltonian_operators.shape[0]),
                n_dis=0 if self._dissipator_operators is None else int(self._dissipator_operators.shape[0]))
        return self._fused

    def fused_tables(self, ham_table, dis_table, device):
        """(M1 table, M2^T table, gamma table) on the device for T times; ham_table (T, Kh) / dis_table (T, Kd) are real
        host arrays or device tensors (None when the collection has no such operators)."""
        f = self.fused_operands()
        parts = []
        T = None
        for tab, cnt in ((ham_table, f["n_ham"]), (dis_table, f["n_dis"])):
            if cnt:
                if tab is None:
                    raise QiskitError("coefficients required for a collection with operators.")
                t = asreal(tab, device).reshape(-1, cnt)
                T = t.shape[0]
                parts.append(t)
        T = 1 if T is None else T
        coeff = torch.cat(parts, dim=1).contiguous() if parts else None
        n = self.dim
        m1 = _abi.generator(n, f["ops1"], f["stat1"], coeff, None, None, layout=_abi.LAYOUT_PACKED)
        m2t = _abi.generator(n, f["ops2t"], f["stat2t"], coeff, None, None, layout=_abi.LAYOUT_PACKED)
        if coeff is None and T > 1:
            m1, m2t = m1.expand(T, -1).contiguous(), m2t.expand(T, -1).contiguous()
        gamma = None
        if f["n_dis"]:
            ones = torch.ones((T, f["n_static"]), dtype=RDTYPE, device=coeff.device)
            gamma = torch.cat([ones, parts[-1]], dim=1).contiguous()
        return m1, m2t, gamma

    def evaluate_rhs(self, ham_coefficients, dis_coefficients, y):
        y = asarray(y)
        single = y.ndim == 2
        rho = y.unsqueeze(0) if single else y
        rho = rho.contiguous()
        has_ham = self._ham is not None
        has_dis = self._static_dissipators is not None or self._dissipator_operators is not None
        if not has_ham and not has_dis:
            raise QiskitError(
                "LindbladCollection with None for static_hamiltonian, hamiltonian_operators, "
                "static_dissipators, and dissipator_operators, cannot evaluate rhs."
            )
        n = self.dim
        if rho.ndim != 3 or tuple(rho.shape[-2:]) != (n, n):
            raise QiskitError(f"state must be (n, n) or (l, n, n) with n = {n}, got {tuple(y.shape)}.")
        if _abi.lindblad_supported(n):
            # fused: two generator launches (M1, M2^T at this time) + ONE kernel for the whole batch; coefficients never
            # come back to the host
            f = self.fused_operands()
            m1, m2t, gamma = self.fused_tables(ham_coefficients, dis_coefficients, rho.device)
            out = _abi.lindblad_rhs(n, m1[0], m2t[0], f["diss"], None if gamma is None else gamma[0].contiguous(), None, 0.0, rho)
            return out[0].contiguous() if single else out
        return self._evaluate_rhs_gemm(ham_coefficients, dis_coefficients, rho, single)

    def _evaluate_rhs_gemm(self, ham_coefficients, dis_coefficients, rho, single):
        """n > 32: the same sum as a sequence of DMMA GEMMs with the batch folded into the free dimension."""
        B = None
        if self._ham is not None:
            B = -1j * self._ham.evaluate(ham_coefficients)
        has_dis = self._static_dissipators is not None or self._dissipator_operators is not None
        if not has_dis:
            out = self._left(B, rho) - self._right(rho, B)
            return out[0].contiguous() if single else out.contiguous()
        if self._dissipator_operators is None:
            A = self._static_product_sum
        else:
            A = self._dis_products.evaluate(dis_coefficients)
        if B is not None:
            out = self._left((B + A).contiguous(), rho) + self._right(rho, (A - B).contiguous())
        else:
            out = self._left(A, rho) + self._right(rho, A)
        if self._static_dissipators is not None:
            for D, Dadj in zip(self._static_dissipators, self._static_adj):
                out = out + self._left(D.contiguous(), self._right(rho, Dadj.contiguous()))
        if self._dissipator_operators is not None:
            g = asreal(dis_coefficients, rho.device).reshape(-1)  # stays on the device: no host round trip
            for j, (L, Ladj) in enumerate(zip(self._dissipator_operators, self._dis_adj)):
                out = out + g[j] * self._left(L.contiguous(), self._right(rho, Ladj.contiguous()))
        return out[0].contiguous() if single else out.contiguous()

    def __call__(self, ham_coefficients, dis_coefficients, y=None):
        if y is None:
            return self.evaluate(ham_coefficients, dis_coefficients)
        return self.evaluate_rhs(ham_coefficients, dis_coefficients, y)
