"""LindbladModel: Lindblad master equation, vectorised (one (n^2, n^2) generator acting on
column-stacked rho -- the LMDE form the fused steppers consume) or not.

Mirror of the reference's ``models/lindblad_model.py`` (constructor, ``from_hamiltonian``,
properties, ``evaluate`` / ``evaluate_rhs`` / ``evaluate_hamiltonian``).
"""

from __future__ import annotations

from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _abi
from ..arrays import asarray, asreal
from ..exceptions import QiskitError
from ..signals import Signal, SignalList
from .generator_model import (BaseGeneratorModel, _operators_into_frame_basis, _static_operator_into_frame_basis,
                              is_hermitian)
from .hamiltonian_model import HamiltonianModel
from .operator_collections import (LindbladCollection, OperatorCollection, VectorizedLindbladCollection,
                                   _as_columns)
from .rotating_frame import RotatingFrame


def _stack(ops):
    ops = asarray(ops)
    if ops is not None and ops.ndim == 2:
        ops = ops.unsqueeze(0).contiguous()
    return ops


class LindbladModel(BaseGeneratorModel):
    def __init__(self, static_hamiltonian=None, hamiltonian_operators=None, hamiltonian_signals=None,
                 static_dissipators=None, dissipator_operators=None, dissipator_signals=None,
                 rotating_frame=None, in_frame_basis: bool = False, array_library: Optional[str] = None,
                 vectorized: bool = False, validate: bool = True):
        if (static_hamiltonian is None and hamiltonian_operators is None and static_dissipators is None
                and dissipator_operators is None):
            raise QiskitError(
                f"{type(self).__name__} requires at least one of static_hamiltonian hamiltonian_operators, "
                "static_dissipators, or dissipator_operators to be specified at construction."
            )
        static_hamiltonian = asarray(static_hamiltonian)
        hamiltonian_operators = _stack(hamiltonian_operators)
        static_dissipators = _stack(static_dissipators)
        dissipator_operators = _stack(dissipator_operators)
        if validate:
            if static_hamiltonian is not None and not is_hermitian(static_hamiltonian):
                raise QiskitError("LinbladModel static_hamiltonian must be Hermitian.")
            if hamiltonian_operators is not None and any(not is_hermitian(op) for op in hamiltonian_operators):
                raise QiskitError("LindbladModel hamiltonian_operators must be Hermitian.")

        self._vectorized = vectorized
        self._rotating_frame = RotatingFrame(rotating_frame)
        self._in_frame_basis = in_frame_basis

        # static Hamiltonian: -i fold, subtract the frame, unfold (lindblad_model.py:173-181)
        stat = None if static_hamiltonian is None else -1j * static_hamiltonian
        stat = _static_operator_into_frame_basis(stat, self._rotating_frame)
        if stat is not None:
            stat = (1j * stat).contiguous()
        kwargs = dict(
            static_hamiltonian=stat,
            hamiltonian_operators=_operators_into_frame_basis(hamiltonian_operators, self._rotating_frame),
            static_dissipators=_operators_into_frame_basis(static_dissipators, self._rotating_frame),
            dissipator_operators=_operators_into_frame_basis(dissipator_operators, self._rotating_frame),
            array_library=array_library,
        )
        self._operator_collection = VectorizedLindbladCollection(**kwargs) if vectorized else LindbladCollection(**kwargs)
        self._hamiltonian_signals = None
        self._dissipator_signals = None
        self.signals = (hamiltonian_signals, dissipator_signals)
        super().__init__(array_library=array_library)

    @classmethod
    def from_hamiltonian(cls, hamiltonian: HamiltonianModel, static_dissipators=None, dissipator_operators=None,
                         dissipator_signals=None, array_library: Optional[str] = None, vectorized: bool = False):
        """Build from a HamiltonianModel, keeping its frame and signals (lindblad_model.py:213-259)."""
        keep = hamiltonian.in_frame_basis
        hamiltonian.in_frame_basis = False
        static_hamiltonian = hamiltonian.static_operator
        hamiltonian_operators = hamiltonian.operators
        hamiltonian.in_frame_basis = keep
        return cls(static_hamiltonian=static_hamiltonian, hamiltonian_operators=hamiltonian_operators,
                   hamiltonian_signals=hamiltonian.signals, static_dissipators=static_dissipators,
                   dissipator_operators=dissipator_operators, dissipator_signals=dissipator_signals,
                   rotating_frame=hamiltonian.rotating_frame, in_frame_basis=hamiltonian.in_frame_basis,
                   array_library=array_library, vectorized=vectorized)

    # -- properties -----------------------------------------------------------------------------
    @property
    def dim(self) -> int:
        c = self._operator_collection
        for ops in (c.static_hamiltonian, c.hamiltonian_operators, c.static_dissipators, c.dissipator_operators):
            if ops is not None:
                return int(ops.shape[-1])
        raise QiskitError("empty LindbladModel")

    @property
    def signals(self) -> Tuple[Optional[SignalList], Optional[SignalList]]:
        return (self._hamiltonian_signals, self._dissipator_signals)

    @signals.setter
    def signals(self, new_signals):
        ham, dis = new_signals
        self._hamiltonian_signals = self._check_signals(ham, self._operator_collection.hamiltonian_operators,
                                                        "Hamiltonian", "hamiltonian_operators")
        self._dissipator_signals = self._check_signals(dis, self._operator_collection.dissipator_operators,
                                                       "Dissipator", "dissipator_operators")

    @staticmethod
    def _check_signals(signals, operators, label: str, opname: str):
        if signals is None:
            return None
        if operators is None:
            raise QiskitError(f"{label} signals must be None if {opname} is None.")
        if isinstance(signals, list):
            signals = SignalList(signals)
        if not isinstance(signals, SignalList):
            raise QiskitError(f"{label} signals specified in unaccepted format.")
        if len(signals) != operators.shape[0]:
            raise QiskitError(f"{label} signals need to have the same length as {label.lower()} operators.")
        return signals

    @property
    def in_frame_basis(self) -> bool:
        return self._in_frame_basis

    @in_frame_basis.setter
    def in_frame_basis(self, value: bool):
        self._in_frame_basis = value

    def _out(self, ops):
        if ops is None:
            return None
        return ops if self.in_frame_basis else self.rotating_frame.operator_out_of_frame_basis(ops)

    static_hamiltonian = property(lambda self: self._out(self._operator_collection.static_hamiltonian))
    hamiltonian_operators = property(lambda self: self._out(self._operator_collection.hamiltonian_operators))
    static_dissipators = property(lambda self: self._out(self._operator_collection.static_dissipators))
    dissipator_operators = property(lambda self: self._out(self._operator_collection.dissipator_operators))

    @property
    def vectorized(self) -> bool:
        return self._vectorized

    @property
    def rotating_frame(self) -> RotatingFrame:
        return self._rotating_frame

    # -- what the fused steppers read (vectorised only) -------------------------------------------
    def _frame_freqs(self):
        return self._rotating_frame.vectorized_frame_freqs

    def _collection(self) -> OperatorCollection:
        if not self._vectorized:
            raise QiskitError("only a vectorized LindbladModel has a single linear generator.")
        return self._operator_collection._operator_collection

    def _signal_table(self, times: np.ndarray) -> Optional[np.ndarray]:
        self._require_signals()
        parts = []
        if self._operator_collection.hamiltonian_operators is not None:
            parts.append(self._hamiltonian_signals.table(times))
        if self._operator_collection.dissipator_operators is not None:
            parts.append(self._dissipator_signals.table(times))
        if not parts:
            return None
        return np.ascontiguousarray(np.concatenate(parts, axis=-1))

    def _signal_table_parts(self, times: np.ndarray):
        """(Hamiltonian signal table (T, Kh) or None, dissipator signal table (T, Kd) or None) on a time grid."""
        self._require_signals()
        c = self._operator_collection
        ham = None if c.hamiltonian_operators is None else np.ascontiguousarray(self._hamiltonian_signals.table(times))
        dis = None if c.dissipator_operators is None else np.ascontiguousarray(self._dissipator_signals.table(times))
        return ham, dis

    def _require_signals(self):
        if self._hamiltonian_signals is None and self._operator_collection.hamiltonian_operators is not None:
            raise QiskitError(
                f"{type(self).__name__} with non-empty hamiltonian operators cannot be evaluated without "
                "hamiltonian signals."
            )
        if self._dissipator_signals is None and self._operator_collection.dissipator_operators is not None:
            raise QiskitError(
                f"{type(self).__name__} with non-empty dissipator operators cannot be evaluated without "
                "dissipator signals."
            )

    def _sig_vals(self, time):
        self._require_signals()
        h = None if self._hamiltonian_signals is None else np.asarray(self._hamiltonian_signals(time), dtype=float)
        d = None if self._dissipator_signals is None else np.asarray(self._dissipator_signals(time), dtype=float)
        return h, d

    # -- evaluation ---------------------------------------------------------------------------
    def evaluate_hamiltonian(self, time: float):
        """H(t) in the frame (lindblad_model.py:415-434)."""
        h = None if self._hamiltonian_signals is None else np.asarray(self._hamiltonian_signals(time), dtype=float)
        ham = self._operator_collection.evaluate_hamiltonian(h)
        if self.rotating_frame.frame_diag is not None:
            ham = self.rotating_frame.operator_into_frame(time, ham, operator_in_frame_basis=True,
                                                          return_in_frame_basis=self._in_frame_basis)
        return ham

    def evaluate(self, time: float):
        """Vectorised generator in the frame (lindblad_model.py:436-475)."""
        h, d = self._sig_vals(time)
        if not self._vectorized:
            raise NotImplementedError("Non-vectorized Lindblad models cannot be represented without a given state.")
        coll = self._collection()
        n2 = coll.dim
        table = self._signal_table(np.array([float(time)]))
        dev = (coll.operators if coll.operators is not None else coll.static_operator).device
        coeff = None if table is None else asreal(table, dev)
        mu = self._frame_freqs()
        times = None if mu is None else torch.tensor([float(time)], dtype=torch.float64, device=dev)
        out = _abi.generator(n2, coll.operators, coll.static_operator, coeff, mu, times).reshape(n2, n2)
        if not self._in_frame_basis and self.rotating_frame.frame_basis is not None:
            VU, VUd = self.rotating_frame.vectorized_frame_basis, self.rotating_frame.vectorized_frame_basis_adjoint
            out = _abi.zgemm(VU, _abi.zgemm(out, VUd))
        return out

    def evaluate_rhs(self, time: float, y):
        """Lindblad RHS at (t, rho) (lindblad_model.py:477-538)."""
        h, d = self._sig_vals(time)
        rf = self.rotating_frame
        if self._vectorized:
            coll = self._collection()
            y2, restore = _as_columns(asarray(y))
            if y2.shape[0] != coll.dim:
                raise QiskitError(f"state has leading dimension {y2.shape[0]}, vectorized dimension is {coll.dim}.")
            table = self._signal_table(np.array([float(time)]))
            coeff = None if table is None else asreal(table.reshape(-1), y2.device)
            change_basis = (not self._in_frame_basis) and rf.frame_basis is not None
            if change_basis:
                y2 = _abi.zgemm(rf.vectorized_frame_basis_adjoint, y2)
            out = _abi.rhs(coll.dim, coll.operators, coll.static_operator, coeff, self._frame_freqs(), float(time), y2)
            if change_basis:
                out = _abi.zgemm(rf.vectorized_frame_basis, out)
            return restore(out)
        # non-vectorised: rho (n,n) or (l,n,n)
        rho = asarray(y)
        coll = self._operator_collection
        n = self.dim
        if _abi.lindblad_supported(n) and rho.ndim in (2, 3) and tuple(rho.shape[-2:]) == (n, n):
            # one fused launch for the whole batch: frame phases, both one-sided products and every dissipator term
            # (qdb_lindblad_rhs_c128); M1(t), M2(t)^T come from two generator launches, coefficients stay on the device
            single = rho.ndim == 2
            r3 = (rho.unsqueeze(0) if single else rho).contiguous()
            change_basis = (not self._in_frame_basis) and rf.frame_basis is not None
            if change_basis:  # U^dag rho U through the DMMA GEMM (batch folded into the free dimension)
                r3 = coll._left(rf.frame_basis_adjoint.contiguous(), coll._right(r3, rf.frame_basis.contiguous())).contiguous()
            f = coll.fused_operands()
            m1, m2t, gamma = coll.fused_tables(h, d, r3.device)
            out = _abi.lindblad_rhs(n, m1[0], m2t[0], f["diss"], None if gamma is None else gamma[0].contiguous(),
                                    rf.frame_freqs, float(time), r3)
            if change_basis:
                out = coll._left(rf.frame_basis.contiguous(), coll._right(out, rf.frame_basis_adjoint.contiguous())).contiguous()
            return out[0].contiguous() if single else out
        if rf.frame_diag is not None:
            rho = rf.operator_out_of_frame(time, rho, operator_in_frame_basis=self._in_frame_basis,
                                           return_in_frame_basis=True)
            out = self._operator_collection.evaluate_rhs(h, d, rho)
            return rf.operator_into_frame(time, out, operator_in_frame_basis=True,
                                          return_in_frame_basis=self._in_frame_basis)
        return self._operator_collection.evaluate_rhs(h, d, rho)
