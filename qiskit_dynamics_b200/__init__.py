"""qiskit_dynamics_b200 -- B200-native (sm_100a) time-evolution hot path behind the
qiskit-dynamics Solver / solve_lmde / HamiltonianModel / LindbladModel / RotatingFrame surface.

Operators, signals tables and state batches live in HBM as torch.complex128 tensors; all hot-path
arithmetic runs in hand-written CUDA (csrc/, C-ABI in include/qdb.h) reached through ctypes.
There is no CPU fallback and no multi-backend dispatch.
"""

__version__ = "0.1.0"
