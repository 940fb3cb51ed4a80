"""HamiltonianModel lives beside GeneratorModel; this module keeps the reference's module path
(models/hamiltonian_model.py)."""
from .generator_model import HamiltonianModel, is_hermitian  # noqa: F401
