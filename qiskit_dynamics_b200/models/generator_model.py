"""GeneratorModel / HamiltonianModel: signals + operator collection + rotating frame.

Mirror of the reference's ``models/generator_model.py`` and ``models/hamiltonian_model.py``
(same constructor arguments, properties, ``evaluate`` / ``evaluate_rhs`` / ``__call__``).
``evaluate_rhs`` is ONE call into the fused CUDA RHS (frame pre-phase, operator sum, GEMM, frame
post-phase); ``evaluate`` is one call into the generator kernel.  Basis changes in and out of the
frame basis (only when ``in_frame_basis`` is False) are DMMA GEMMs.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Union

import numpy as np
import torch

from .. import _abi
from ..arrays import asarray, asreal, stage_to_device
from ..exceptions import QiskitError
from ..signals import Signal, SignalList, compile_signal_program, mutation_epoch
from .operator_collections import OperatorCollection, _as_columns
from .rotating_frame import RotatingFrame


class BaseGeneratorModel(ABC):
    """Interface of a linear ODE  dy/dt = Lambda(t, y)  (generator_model.py:41-105)."""

    def __init__(self, array_library: Optional[str] = None):
        self._array_library = array_library

    @property
    @abstractmethod
    def dim(self) -> int:
        ...

    @property
    @abstractmethod
    def rotating_frame(self) -> RotatingFrame:
        ...

    @property
    @abstractmethod
    def in_frame_basis(self) -> bool:
        ...

    @property
    def array_library(self):
        return self._array_library

    @abstractmethod
    def evaluate(self, time: float):
        ...

    @abstractmethod
    def evaluate_rhs(self, time: float, y):
        ...

    def __call__(self, time: float, y=None):
        return self.evaluate(time) if y is None else self.evaluate_rhs(time, y)


def _static_operator_into_frame_basis(static_operator, rotating_frame: RotatingFrame):
    """U^dag G_d U - diag(d), or diag(-d) when only a frame is given (generator_model.py:319-340)."""
    if static_operator is None:
        if rotating_frame.frame_operator is None:
            return None
        return torch.diag(-rotating_frame.frame_diag).contiguous()
    static_operator = asarray(static_operator)
    if rotating_frame.frame_operator is None:
        return static_operator
    out = rotating_frame.operator_into_frame_basis(static_operator)
    return (out - torch.diag(rotating_frame.frame_diag)).contiguous()


def _operators_into_frame_basis(operators, rotating_frame: RotatingFrame):
    """U^dag G_j U (generator_model.py:343-365)."""
    if operators is None:
        return None
    return rotating_frame.operator_into_frame_basis(asarray(operators))


def _signal_values(signals: Optional[SignalList], time) -> Optional[np.ndarray]:
    return None if signals is None else np.asarray(signals(time), dtype=np.float64)


class GeneratorModel(BaseGeneratorModel):
    """G(t) = G_d + sum_i s_i(t) G_i, optionally in a rotating frame (generator_model.py:108-316)."""

    def __init__(self, static_operator=None, operators=None, signals=None, rotating_frame=None,
                 in_frame_basis: bool = False, array_library: Optional[str] = None):
        if static_operator is None and operators is None:
            raise QiskitError(
                f"{type(self).__name__} requires at least one of static_operator or operators to be "
                "specified at construction."
            )
        self._rotating_frame = RotatingFrame(rotating_frame)
        self._in_frame_basis = in_frame_basis
        static_fb = _static_operator_into_frame_basis(static_operator, self._rotating_frame)
        ops_fb = _operators_into_frame_basis(operators, self._rotating_frame)
        self._operator_collection = OperatorCollection(static_operator=static_fb, operators=ops_fb,
                                                       array_library=array_library)
        self._signals = None
        self.signals = signals
        super().__init__(array_library=array_library)

    # -- properties -----------------------------------------------------------------------------
    @property
    def dim(self) -> int:
        return self._operator_collection.dim

    @property
    def rotating_frame(self) -> RotatingFrame:
        return self._rotating_frame

    @property
    def in_frame_basis(self) -> bool:
        return self._in_frame_basis

    @in_frame_basis.setter
    def in_frame_basis(self, value: bool):
        self._in_frame_basis = value

    @property
    def static_operator(self):
        st = self._operator_collection.static_operator
        if st is None:
            return None
        return st if self.in_frame_basis else self.rotating_frame.operator_out_of_frame_basis(st)

    @property
    def operators(self):
        ops = self._operator_collection.operators
        if ops is None:
            return None
        return ops if self.in_frame_basis else self.rotating_frame.operator_out_of_frame_basis(ops)

    @property
    def signals(self) -> Optional[SignalList]:
        return self._signals

    @signals.setter
    def signals(self, signals: Union[SignalList, List[Signal], None]):
        self._signal_program = False  # not compiled yet (None = compiled, not device-evaluable)
        if signals is None:
            self._signals = None
            return
        if self._operator_collection.operators is None:
            raise QiskitError("Signals must be None if operators is None.")
        if isinstance(signals, list):
            signals = SignalList(signals)
        if not isinstance(signals, SignalList):
            raise QiskitError("Signals specified in unaccepted format.")
        if len(signals) != self._operator_collection.num_operators:
            raise QiskitError("Signals needs to have the same length as operators.")
        self._signals = signals

    # -- what the fused steppers read -----------------------------------------------------------
    def _frame_freqs(self):
        return self._rotating_frame.frame_freqs

    def _collection(self) -> OperatorCollection:
        return self._operator_collection

    def _signal_table(self, times: np.ndarray) -> Optional[np.ndarray]:
        """(T, K) float64 table of signal values on a time grid (one vectorised host call)."""
        if self._operator_collection.operators is None:
            return None
        self._require_signals()
        return self._signals.table(times)

    def _program(self):
        """Device program of the current signals (compiled once per assignment), or None when a term is an
        arbitrary Python envelope."""
        if getattr(self, "_signal_program", False) is False or getattr(self, "_signal_program_epoch", -1) != mutation_epoch():
            self._signal_program = compile_signal_program(self._signals) if self._signals is not None else None
            self._signal_program_epoch = mutation_epoch()
        return self._signal_program

    def _device_coefficients(self, time, device):
        """(1, K) signal values at one time, evaluated on the device when every term is a sampled or
        constant-envelope signal (no host NumPy, no host-to-device copy); else None."""
        prog = self._program()
        if prog is None or np.ndim(time) != 0:
            return None
        return prog.table(float(time), device)

    def _signal_table_device(self, times: np.ndarray, device):
        """(T, K) table on a time grid built by the device signal kernel (row f3), or None: the host then falls
        back to :meth:`_signal_table`.  The grid goes up through a pinned staging buffer, asynchronously, so that
        the generator and stepper launches are enqueued while the state batch is still in flight."""
        if self._operator_collection.operators is None:
            return None
        prog = self._program()
        if prog is None:
            return None
        return prog.table(stage_to_device(times, device), device)

    def _require_signals(self):
        if self._signals is None and self._operator_collection.operators is not None:
            raise QiskitError(f"{type(self).__name__} with non-empty operators must be evaluated signals.")

    # -- evaluation ---------------------------------------------------------------------------
    def evaluate(self, time: float):
        """G(t) in the frame: (G_d + sum s_j G_j) .* outer(conj e, e) (generator_model.py:256-279)."""
        self._require_signals()
        coll = self._operator_collection
        n = coll.dim
        sig = _signal_values(self._signals, time)
        dev = (coll.operators if coll.operators is not None else coll.static_operator).device
        coeff = None if sig is None else asreal(sig.reshape(1, -1), dev)
        mu = self._frame_freqs()
        times = None if mu is None else torch.tensor([float(time)], dtype=torch.float64, device=dev)
        out = _abi.generator(n, coll.operators, coll.static_operator, coeff, mu, times).reshape(n, n)
        if not self._in_frame_basis:
            out = self.rotating_frame.operator_out_of_frame_basis(out)
        return out

    def evaluate_rhs(self, time: float, y):
        """G(t) y with the frame rotations fused in (generator_model.py:281-316)."""
        self._require_signals()
        coll = self._operator_collection
        y2, restore = _as_columns(asarray(y))
        if y2.shape[0] != coll.dim:
            raise QiskitError(f"state has leading dimension {y2.shape[0]}, model dimension is {coll.dim}.")
        coeff = self._device_coefficients(time, y2.device)
        if coeff is None:
            sig = _signal_values(self._signals, time)
            coeff = None if sig is None else asreal(sig, y2.device)
        if not self._in_frame_basis:
            y2 = self.rotating_frame.state_into_frame_basis(y2)
        out = _abi.rhs(coll.dim, coll.operators, coll.static_operator, coeff, self._frame_freqs(), float(time), y2)
        if not self._in_frame_basis:
            out = self.rotating_frame.state_out_of_frame_basis(out)
        return restore(out)


def is_hermitian(operator, tol: float = 1e-10) -> bool:
    """|| A^dag - A || < tol (hamiltonian_model.py:153-178)."""
    op = asarray(operator)
    return bool(torch.linalg.norm(op.conj().transpose(-1, -2) - op) < tol)


class HamiltonianModel(GeneratorModel):
    """H(t) = H_d + sum_j s_j(t) H_j with generator -iH (hamiltonian_model.py:32-150)."""

    def __init__(self, static_operator=None, operators=None, signals=None, rotating_frame=None,
                 in_frame_basis: bool = False, array_library: Optional[str] = None, validate: bool = True):
        if static_operator is not None:
            static_operator = asarray(static_operator)
            if validate and not is_hermitian(static_operator):
                raise QiskitError("HamiltonianModel static_operator must be Hermitian.")
            static_operator = -1j * static_operator
        if operators is not None:
            operators = asarray(operators)
            if operators.ndim == 2:
                operators = operators.unsqueeze(0)
            if validate and any(not is_hermitian(op) for op in operators):
                raise QiskitError("HamiltonianModel operators must be Hermitian.")
            operators = -1j * operators
        super().__init__(static_operator=static_operator, operators=operators, signals=signals,
                         rotating_frame=rotating_frame, in_frame_basis=in_frame_basis,
                         array_library=array_library)

    @property
    def static_operator(self):
        st = self._operator_collection.static_operator
        if st is None:
            return None
        if self.in_frame_basis:
            return st
        return 1j * self.rotating_frame.operator_out_of_frame_basis(st)

    @property
    def operators(self):
        ops = self._operator_collection.operators
        if ops is None:
            return None
        if not self.in_frame_basis:
            ops = self.rotating_frame.operator_out_of_frame_basis(ops)
        return 1j * ops
