"""qiskit_dynamics_b200 -- B200-native (sm_100a) time-evolution hot path behind the
qiskit-dynamics Solver / solve_lmde / HamiltonianModel / LindbladModel / RotatingFrame surface.

Operators, signal tables and state batches live in HBM as torch.complex128 tensors; all hot-path
arithmetic runs in hand-written CUDA (csrc/, C-ABI in include/qdb.h) reached through ctypes.
There is no CPU fallback and no multi-backend dispatch.
"""

__version__ = "0.1.0"

from .exceptions import QiskitError
from .arrays import asarray, default_device, set_default_device, to_numpy
from .signals import Signal, DiscreteSignal, SignalSum, DiscreteSignalSum, SignalList
from .models import (RotatingFrame, GeneratorModel, HamiltonianModel, LindbladModel, OperatorCollection,
                     LindbladCollection, VectorizedLindbladCollection)
from .solvers import Solver, solve_lmde, solve_ode
from . import distributed
from . import measurement
from .measurement import FinalStateMeasurement

__all__ = ["FinalStateMeasurement", "QiskitError", "Signal", "DiscreteSignal", "SignalSum", "DiscreteSignalSum", "SignalList",
           "RotatingFrame", "GeneratorModel", "HamiltonianModel", "LindbladModel", "OperatorCollection",
           "LindbladCollection", "VectorizedLindbladCollection", "Solver", "solve_lmde", "solve_ode",
           "asarray", "default_device", "set_default_device", "to_numpy", "distributed"]
