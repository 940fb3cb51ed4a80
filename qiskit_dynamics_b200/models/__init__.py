"""Models: signals + operators + rotating frame (mirror of qiskit_dynamics.models)."""
from .rotating_frame import RotatingFrame
from .generator_model import BaseGeneratorModel, GeneratorModel, HamiltonianModel, is_hermitian
from .lindblad_model import LindbladModel
from .operator_collections import (OperatorCollection, LindbladCollection, VectorizedLindbladCollection,
                                   vec_commutator, vec_dissipator)

__all__ = ["RotatingFrame", "BaseGeneratorModel", "GeneratorModel", "HamiltonianModel", "LindbladModel",
           "OperatorCollection", "LindbladCollection", "VectorizedLindbladCollection", "vec_commutator",
           "vec_dissipator", "is_hermitian"]
