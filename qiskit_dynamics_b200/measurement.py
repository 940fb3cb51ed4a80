"""Final-state measurement post-processing on the device.

Mirror of the part of ``DynamicsBackend`` that follows the solve (backend/dynamics_backend.py:846-866,
backend/backend_utils.py:31-147): final states out of the rotating frame, into the dressed basis, normalised,
and reduced to memory-slot outcome probabilities -- for a whole batch of state columns at once:

    y_meas = V^dagger U diag(exp(d t)) U^dagger y        two qdb_zgemm_c128 calls (phases in `pre`)
    P[o, b] = sum_{i -> o} |y_meas[i, b]|^2 / ||y_meas[:, b]||^2        qdb_outcome_probabilities_f64

The (n_out, B) probability table is the "final observables" object that the multi-GPU path all-gathers
(:func:`qiskit_dynamics_b200.distributed.all_gather_columns`).  Set-up (eigendecomposition, the basis-state ->
outcome map) is host NumPy, once per backend.  Sampling counts from the probabilities stays on the host
(NumPy generator, as in the reference).  The qiskit ``Statevector`` / ``Result`` wrappers are outside this build.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi
from .arrays import asarray, asreal
from .exceptions import QiskitError
from .models import HamiltonianModel


def get_dressed_state_decomposition(operator, rtol: float = 1e-8, atol: float = 1e-5) -> Tuple[np.ndarray, np.ndarray]:
    """Eigenvalues / eigenvectors of a nearly diagonal Hermitian operator, sorted by overlap with the elementary
    basis (backend/backend_utils.py:31-80)."""
    op = np.array(operator.detach().cpu().numpy() if isinstance(operator, torch.Tensor) else operator)
    if not np.allclose(op, op.conj().T, rtol=rtol, atol=atol):
        raise QiskitError("_get_dressed_state_decomposition received non-Hermitian operator.")
    evals, evecs = np.linalg.eigh(op)
    dressed_evals = np.zeros_like(evals)
    dressed_states = np.zeros_like(evecs)
    found = []
    for eigval, evec in zip(evals, evecs.transpose()):
        position = int(np.argmax(np.abs(evec)))
        if position in found:
            raise QiskitError("Dressed-state sorting failed due to non-unique np.argmax(np.abs(evec)) for eigenvectors.")
        found.append(position)
        dressed_states[:, position] = evec
        dressed_evals[position] = eigval
    return dressed_evals, dressed_states


def get_lab_frame_static_hamiltonian(model) -> np.ndarray:
    """Static Hamiltonian in the lab frame and standard basis (backend/backend_utils.py:83-103)."""
    static = model.static_operator if isinstance(model, HamiltonianModel) else model.static_hamiltonian
    out = 1j * model.rotating_frame.generator_out_of_frame(t=0.0, operator=-1j * asarray(static))
    return out.detach().cpu().numpy()


def memory_slot_outcome_map(subsystem_dims: Sequence[int], measurement_subsystems: Sequence[int],
                            memory_slot_indices: Sequence[int], num_memory_slots: Optional[int] = None,
                            max_outcome_level: Optional[int] = None) -> Tuple[List[str], np.ndarray]:
    """(labels, outcome_of): the memory-slot outcome string of every basis state.

    Basis index i = sum_s level_s * prod_{r<s} dims[r] (subsystem 0 least significant, the qiskit convention of
    Statevector.probabilities_dict); measured subsystem k writes min(level, max_outcome_level) to memory slot
    memory_slot_indices[k], unused slots read "0" (backend/backend_utils.py:106-147).  Labels are sorted."""
    dims = [int(d) for d in subsystem_dims]
    n = int(np.prod(dims))
    num_memory_slots = num_memory_slots or (max(memory_slot_indices) + 1)
    keys = []
    for i in range(n):
        rem, levels = i, []
        for d in dims:
            levels.append(rem % d)
            rem //= d
        result = ["0"] * num_memory_slots
        for slot, sub in zip(memory_slot_indices, measurement_subsystems):
            level = levels[sub]
            if max_outcome_level and level > max_outcome_level:
                level = max_outcome_level
            result[-(slot + 1)] = str(level)
        keys.append("".join(result))
    labels = sorted(set(keys))
    index = {k: j for j, k in enumerate(labels)}
    return labels, np.asarray([index[k] for k in keys], dtype=np.int32)


class FinalStateMeasurement:
    """Prepared measurement of batches of final states of one model.

    Args mirror the ``DynamicsBackend`` options that enter the post-processing: ``subsystem_dims``,
    ``measurement_subsystems`` / ``memory_slot_indices`` / ``num_memory_slots`` (the measurement specification of
    an experiment), ``max_outcome_level``, ``normalize_states``; ``dressed_states`` defaults to the decomposition
    of the model's lab-frame static Hamiltonian (dynamics_backend.py: ``_dressed_states``).
    """

    def __init__(self, model, subsystem_dims: Sequence[int], measurement_subsystems: Sequence[int],
                 memory_slot_indices: Optional[Sequence[int]] = None, num_memory_slots: Optional[int] = None,
                 max_outcome_level: Optional[int] = 1, normalize_states: bool = True, dressed_states=None):
        self.model = model
        n = model.dim
        if int(np.prod(subsystem_dims)) != n:
            raise QiskitError("subsystem_dims do not multiply to the model dimension.")
        if memory_slot_indices is None:
            memory_slot_indices = list(range(len(measurement_subsystems)))
        if dressed_states is None:
            _, dressed_states = get_dressed_state_decomposition(get_lab_frame_static_hamiltonian(model))
        self.dressed_states = np.asarray(dressed_states)
        self.normalize_states = bool(normalize_states)
        self.labels, outcome_of = memory_slot_outcome_map(subsystem_dims, measurement_subsystems, memory_slot_indices,
                                                          num_memory_slots, max_outcome_level)
        frame = model.rotating_frame
        Vh = asarray(self.dressed_states.conj().T.copy())
        self._device = Vh.device
        self._outcome_of = torch.from_numpy(outcome_of).to(self._device)
        U = frame.frame_basis
        if U is None:
            self._Uh, self._W = None, Vh.contiguous()
        else:
            self._Uh = frame.frame_basis_adjoint.contiguous()
            self._W = _abi.zgemm(Vh.contiguous(), U.contiguous())  # V^dagger U, once
        d = frame.frame_diag
        self._mu = None if d is None else asreal((-d.imag).contiguous(), self._device)  # exp(d t) = exp(-i mu t)

    def measurement_basis_states(self, t: float, y: torch.Tensor) -> torch.Tensor:
        """Columns of y (standard basis, in the rotating frame at time t) in the lab frame and dressed basis."""
        y = asarray(y)
        vec = y.ndim == 1
        Y = y.reshape(-1, 1).contiguous() if vec else y.contiguous()
        if self._Uh is not None:
            Y = _abi.zgemm(self._Uh, Y)
        pre = None
        if self._mu is not None:
            ang = self._mu * float(t)
            pre = torch.complex(torch.cos(ang), -torch.sin(ang)).contiguous()
        out = _abi.zgemm(self._W, Y, pre=pre)
        return out.reshape(-1) if vec else out

    def probabilities(self, t: float, y: torch.Tensor) -> torch.Tensor:
        """(n_out, B) memory-slot outcome probabilities (rows follow ``self.labels``)."""
        Y = self.measurement_basis_states(t, y)
        if Y.ndim == 1:
            Y = Y.reshape(-1, 1).contiguous()
        return _abi.outcome_probabilities(Y, self._outcome_of, len(self.labels), normalize=self.normalize_states)

    def probabilities_dicts(self, t: float, y: torch.Tensor) -> List[Dict[str, float]]:
        """Per column, the dictionary the reference builds (zero-probability outcomes dropped)."""
        P = self.probabilities(t, y).cpu().numpy()
        return [{lab: float(p) for lab, p in zip(self.labels, col) if p != 0.0} for col in P.T]

    def sample_counts(self, t: float, y: torch.Tensor, shots: int, seed: Optional[int] = None) -> List[Dict[str, int]]:
        """Counts per column: NumPy ``Generator.choice`` over the outcome dictionary, like
        _sample_probability_dict + _get_counts_from_samples (backend/backend_utils.py:150-185)."""
        out = []
        for pd in self.probabilities_dicts(t, y):
            rng = np.random.default_rng(seed=seed)
            alphabet, probs = zip(*pd.items())
            probs = np.array(probs)
            if self.normalize_states:
                probs = probs / probs.sum()
            samples = rng.choice(alphabet, size=shots, replace=True, p=probs)
            out.append({str(k): int(v) for k, v in zip(*np.unique(samples, return_counts=True))})
        return out
