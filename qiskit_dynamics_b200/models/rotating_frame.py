"""RotatingFrame: maps into/out of the frame of an anti-Hermitian F = -iH.

Mirror of the reference's ``models/rotating_frame.py`` (same constructor, properties and method
names/flags).  Set-up (``eigh``, Kronecker bases) is one-off torch linear algebra; every method
that touches a *time* or a *state batch* goes through the CUDA C-ABI:

* diagonal phases on states   -> ``qdb_frame_apply_c128``      (rotating_frame.py:255)
* diagonal phases on operators-> ``qdb_generator_c128`` (K = 0)  (rotating_frame.py:350-353)
* basis changes               -> ``qdb_zgemm_c128``            (rotating_frame.py:152,167,195,223)

Inside the solvers none of these are called per step: the fused kernels take ``frame_freqs``
(``mu``, real) and apply e^{-+i mu t} themselves.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _abi
from ..arrays import asarray, CDTYPE
from ..exceptions import QiskitError


def _is_close_to(a: torch.Tensor, b: torch.Tensor, atol: float, rtol: float) -> bool:
    return bool(torch.all((a - b).abs() <= atol + rtol * b.abs()))


def _enforce_anti_herm(mat: torch.Tensor, atol: float = 1e-10, rtol: float = 1e-10) -> torch.Tensor:
    """Hermitian H -> -iH; anti-Hermitian F -> F; otherwise error (rotating_frame.py:585-660)."""
    adj = mat.conj() if mat.ndim == 1 else mat.conj().transpose(-1, -2)
    if _is_close_to(mat, adj, atol, rtol):
        return -1j * mat
    if _is_close_to(mat, -adj, atol, rtol):
        return mat
    raise QiskitError("frame_operator must be either a Hermitian or anti-Hermitian matrix.")


class RotatingFrame:
    def __init__(self, frame_operator, atol: float = 1e-10, rtol: float = 1e-10):
        if isinstance(frame_operator, RotatingFrame):
            frame_operator = frame_operator.frame_operator
        self._frame_operator = frame_operator
        self._frame_basis = None
        self._frame_basis_adjoint = None
        self._frame_diag = None
        self._dim = None
        self._vectorized_frame_basis = None
        self._vectorized_frame_basis_adjoint = None
        self._mu = None
        self._mu_vec = None
        if frame_operator is None:
            return
        F = _enforce_anti_herm(asarray(frame_operator), atol=atol, rtol=rtol)
        if F.ndim == 1:
            self._frame_diag = F.contiguous()
        else:
            # eigh(iF) = eigh(H): real eigenvalues lam, unitary U; frame_diag = -i lam (:103-108)
            lam, U = torch.linalg.eigh(1j * F)
            self._frame_diag = (-1j * lam.to(CDTYPE)).contiguous()
            self._frame_basis = U.contiguous()
            self._frame_basis_adjoint = U.conj().resolve_conj().transpose(0, 1).contiguous()
        self._dim = int(self._frame_diag.shape[0])
        self._mu = (-self._frame_diag.imag).contiguous()  # lam: real frame frequencies

    # -- properties -----------------------------------------------------------------------------
    @property
    def dim(self):
        return self._dim

    @property
    def frame_operator(self):
        return self._frame_operator

    @property
    def frame_diag(self):
        return self._frame_diag

    @property
    def frame_basis(self):
        return self._frame_basis

    @property
    def frame_basis_adjoint(self):
        return self._frame_basis_adjoint

    @property
    def frame_freqs(self) -> Optional[torch.Tensor]:
        """mu = -Im(frame_diag) (real, length dim): what the kernels take; phases are exp(-i mu t)."""
        return self._mu

    @property
    def vectorized_frame_freqs(self) -> Optional[torch.Tensor]:
        """mu for column-stacked density matrices: mu[i + k dim] = lam_i - lam_k."""
        if self._mu is None:
            return None
        if self._mu_vec is None:
            lam = self._mu
            self._mu_vec = (lam[None, :] - lam[:, None]).reshape(-1).contiguous()  # [k, i] -> i + k dim
        return self._mu_vec

    # -- basis changes --------------------------------------------------------------------------
    @staticmethod
    def _left_multiply(M: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """M @ y for y of shape (n,), (n, B) or (T, n, B) through the DMMA GEMM."""
        if y.ndim == 1:
            return _abi.zgemm(M, y.reshape(-1, 1).contiguous()).reshape(-1)
        if y.ndim == 2:
            return _abi.zgemm(M, y.contiguous())
        return torch.stack([RotatingFrame._left_multiply(M, yi) for yi in y])

    def state_into_frame_basis(self, y):
        y = asarray(y)
        if self._frame_basis_adjoint is None:
            return y
        return self._left_multiply(self._frame_basis_adjoint, y)

    def state_out_of_frame_basis(self, y):
        y = asarray(y)
        if self._frame_basis is None:
            return y
        return self._left_multiply(self._frame_basis, y)

    def operator_into_frame_basis(self, op, convert_type: bool = True):
        """U^dag A U for one operator or a stack (set-up time; torch matmul)."""
        if op is None:
            return None
        if isinstance(op, list) and not convert_type:
            return [self.operator_into_frame_basis(x, convert_type=False) for x in op]
        op = asarray(op)
        if self._frame_basis is None:
            return op
        return torch.matmul(self._frame_basis_adjoint, torch.matmul(op, self._frame_basis)).contiguous()

    def operator_out_of_frame_basis(self, op, convert_type: bool = True):
        if op is None:
            return None
        if isinstance(op, list) and not convert_type:
            return [self.operator_out_of_frame_basis(x, convert_type=False) for x in op]
        op = asarray(op)
        if self._frame_basis is None:
            return op
        return torch.matmul(self._frame_basis, torch.matmul(op, self._frame_basis_adjoint)).contiguous()

    # -- states ---------------------------------------------------------------------------------
    def state_into_frame(self, t: float, y, y_in_frame_basis: bool = False, return_in_frame_basis: bool = False):
        """exp(-tF) y (rotating_frame.py:225-261)."""
        y = asarray(y)
        if self._frame_operator is None:
            return y
        out = y if y_in_frame_basis else self.state_into_frame_basis(y)
        shape = out.shape
        out2 = out.reshape(shape[0], -1).contiguous()
        # exp(frame_diag * (-t)) = exp(+i mu t) = conj(p(t))
        out2 = _abi.frame_apply(self._mu, t, out2, conj_phase=True)
        out = out2.reshape(shape)
        return out if return_in_frame_basis else self.state_out_of_frame_basis(out)

    def state_out_of_frame(self, t: float, y, y_in_frame_basis: bool = False, return_in_frame_basis: bool = False):
        """exp(tF) y: into-frame with time reversed (rotating_frame.py:263-284)."""
        return self.state_into_frame(-t, y, y_in_frame_basis, return_in_frame_basis)

    # -- operators ------------------------------------------------------------------------------
    def _phase_operator(self, t: float, op: torch.Tensor) -> torch.Tensor:
        """op .* outer(conj(e), e), e = exp(frame_diag t), for (n,n) or (k,n,n) in the frame basis."""
        n = self._dim
        if op.ndim == 2:
            times = torch.tensor([t], dtype=torch.float64, device=op.device)
            return _abi.generator(n, None, op.contiguous(), None, self._mu, times).reshape(n, n)
        # a stack of operators: one broadcast multiply with the (n, n) phase matrix (set-up / API glue; the solvers and
        # LindbladModel.evaluate_rhs apply these phases inside their kernels and never come here)
        ang = self._mu * float(t)
        e = torch.complex(torch.cos(ang), -torch.sin(ang))  # exp(-i mu t) = exp(frame_diag t)
        return (op * (e.conj()[:, None] * e[None, :])).contiguous()

    def _conjugate_and_add(self, t, operator, op_to_add_in_fb=None, operator_in_frame_basis=False,
                           return_in_frame_basis=False, vectorized_operators=False):
        """exp(-tF) G exp(tF) + B with B added in the frame basis (rotating_frame.py:286-370)."""
        operator = asarray(operator)
        add = asarray(op_to_add_in_fb)
        if vectorized_operators:
            if self._frame_operator is None:
                return operator if add is None else operator + add
            # (dim^2,) or (dim^2, k) column-stacked -> (dim, dim) or (k, dim, dim)
            if operator.ndim == 2:
                operator = operator.transpose(0, 1)
            lead = operator.shape[:-1]
            operator = operator.reshape(lead + (self._dim, self._dim)).transpose(-1, -2)  # order="F"
        if self._frame_operator is None:
            return operator if add is None else operator + add
        out = operator if operator_in_frame_basis else self.operator_into_frame_basis(operator)
        out = self._phase_operator(t, out.contiguous())
        if add is not None:
            out = out + add
        if not return_in_frame_basis:
            out = self.operator_out_of_frame_basis(out)
        if vectorized_operators:
            out = out.transpose(-1, -2).reshape(out.shape[:-2] + (self._dim**2,))
            if out.ndim == 2:
                out = out.transpose(0, 1)
            out = out.contiguous()
        return out

    def operator_into_frame(self, t, operator, operator_in_frame_basis=False, return_in_frame_basis=False,
                            vectorized_operators=False):
        return self._conjugate_and_add(t, operator, operator_in_frame_basis=operator_in_frame_basis,
                                       return_in_frame_basis=return_in_frame_basis,
                                       vectorized_operators=vectorized_operators)

    def operator_out_of_frame(self, t, operator, operator_in_frame_basis=False, return_in_frame_basis=False,
                              vectorized_operators=False):
        return self.operator_into_frame(-t, operator, operator_in_frame_basis=operator_in_frame_basis,
                                        return_in_frame_basis=return_in_frame_basis,
                                        vectorized_operators=vectorized_operators)

    def generator_into_frame(self, t, operator, operator_in_frame_basis=False, return_in_frame_basis=False,
                             vectorized_operators=False):
        """exp(-tF) G exp(tF) - F (rotating_frame.py:438-474)."""
        if self._frame_operator is None:
            return asarray(operator)
        return self._conjugate_and_add(t, operator, op_to_add_in_fb=-torch.diag(self._frame_diag),
                                       operator_in_frame_basis=operator_in_frame_basis,
                                       return_in_frame_basis=return_in_frame_basis,
                                       vectorized_operators=vectorized_operators)

    def generator_out_of_frame(self, t, operator, operator_in_frame_basis=False, return_in_frame_basis=False):
        if self._frame_operator is None:
            return asarray(operator)
        return self._conjugate_and_add(-t, operator, op_to_add_in_fb=torch.diag(self._frame_diag),
                                       operator_in_frame_basis=operator_in_frame_basis,
                                       return_in_frame_basis=return_in_frame_basis)

    # -- vectorised maps ------------------------------------------------------------------------
    @property
    def vectorized_frame_basis(self):
        """conj(U) kron U, lazily (rotating_frame.py:510-521)."""
        if self._frame_basis is None:
            return None
        if self._vectorized_frame_basis is None:
            self._vectorized_frame_basis = torch.kron(self._frame_basis.conj().resolve_conj(), self._frame_basis).contiguous()
            self._vectorized_frame_basis_adjoint = self._vectorized_frame_basis.conj().resolve_conj().transpose(0, 1).contiguous()
        return self._vectorized_frame_basis

    @property
    def vectorized_frame_basis_adjoint(self):
        if self._frame_basis is None:
            return None
        if self._vectorized_frame_basis_adjoint is None:
            self.vectorized_frame_basis  # noqa: B018  (triggers the lazy build)
        return self._vectorized_frame_basis_adjoint

    def vectorized_map_into_frame(self, time, op, operator_in_frame_basis=False, return_in_frame_basis=False):
        """(dim^2, dim^2) superoperator into the frame (rotating_frame.py:537-582)."""
        op = asarray(op)
        if self._frame_diag is None:
            return op
        if not operator_in_frame_basis and self._frame_basis is not None:
            op = _abi.zgemm(self.vectorized_frame_basis_adjoint, _abi.zgemm(op.contiguous(), self.vectorized_frame_basis))
        n2 = self._dim**2
        times = torch.tensor([time], dtype=torch.float64, device=op.device)
        # Hadamard with outer(conj(tau), tau): the generator kernel with the vectorised frequencies
        op = _abi.generator(n2, None, op.contiguous(), None, self.vectorized_frame_freqs, times).reshape(n2, n2)
        if not return_in_frame_basis and self._frame_basis is not None:
            op = _abi.zgemm(self.vectorized_frame_basis, _abi.zgemm(op, self.vectorized_frame_basis_adjoint))
        return op
