"""solve_ode / solve_lmde: the method-string dispatch of the reference
(solvers/solver_functions.py:129-373) with the fixed-step methods served by the fused kernels.

Supported ``method`` strings: ``"RK4"`` (and its JAX alias ``"jax_RK4"``), ``"scipy_expm"`` (aliases
``"jax_expm"``, ``"expm"``; ``magnus_order`` 1, 2, 3) and the time-parallel LMDE solvers
``"jax_RK4_parallel"`` / ``"jax_expm_parallel"`` (propagators of all steps built side by side, then
multiplied; model generators only).  The reference's adaptive SciPy/JAX/diffrax integrators and Lanczos
are a different algorithm family and outside this build: asking for
them raises ``QiskitError``.
"""

from __future__ import annotations

from typing import Callable, Optional, Union

import numpy as np
import torch
from scipy.integrate._ivp.ivp import OdeResult

from .. import _abi
from ..arrays import asarray, wait_pending_copies
from ..exceptions import QiskitError
from ..models import BaseGeneratorModel, GeneratorModel, LindbladModel
from .fixed_step import (RK4_solver, expm_model_solve, lindblad_rk4_model_solve, parallel_model_solve, rk4_model_solve,
                         scipy_expm_solver)

ODE_METHODS = ["RK4", "jax_RK4"]
LMDE_METHODS = ["scipy_expm", "jax_expm", "expm", "jax_RK4_parallel", "jax_expm_parallel"]
PARALLEL_METHODS = {"jax_RK4_parallel": "RK4", "jax_expm_parallel": "expm"}
_REFERENCE_ONLY = ["RK45", "RK23", "BDF", "DOP853", "Radau", "LSODA", "jax_odeint", "lanczos_diag", "jax_lanczos_diag"]


def _unsupported(method, who: str):
    if method in _REFERENCE_ONLY:
        return QiskitError(
            f"Method {method} is not part of the B200 build (fixed-step RK4, scipy_expm and their time-parallel variants only); "
            f"it is not supported by {who}."
        )
    return QiskitError(f"Method {method} not supported by {who}.")


def is_lindblad_model_vectorized(obj) -> bool:
    return isinstance(obj, LindbladModel) and obj.vectorized


def is_lindblad_model_not_vectorized(obj) -> bool:
    return isinstance(obj, LindbladModel) and not obj.vectorized


def _has_linear_generator(model) -> bool:
    """True when the model is a single linear generator acting on columns (fusable)."""
    return isinstance(model, GeneratorModel) or is_lindblad_model_vectorized(model)


def setup_generator_model_rhs_y0_in_frame_basis(generator_model: BaseGeneratorModel, y0):
    """Put y0 into the frame basis and switch the model to frame-basis evaluation
    (solvers/solver_functions.py:376-415).  Returns (generator, rhs, y0, model_was_in_frame_basis)."""
    was = generator_model.in_frame_basis
    rf = generator_model.rotating_frame
    if not was:
        if is_lindblad_model_vectorized(generator_model):
            if rf.frame_basis is not None:
                y0 = rf._left_multiply(rf.vectorized_frame_basis_adjoint, y0)
        elif isinstance(generator_model, LindbladModel):
            y0 = rf.operator_into_frame_basis(y0)
        elif isinstance(generator_model, GeneratorModel):
            y0 = rf.state_into_frame_basis(y0)
    generator_model.in_frame_basis = True
    return (lambda t: generator_model(t)), (lambda t, y: generator_model(t, y)), y0, was


def results_y_out_of_frame_basis(generator_model: BaseGeneratorModel, results_y: torch.Tensor, y0_ndim: int):
    """Results (T, *y0.shape) back out of the frame basis (solvers/solver_functions.py:418-450)."""
    rf = generator_model.rotating_frame
    if is_lindblad_model_vectorized(generator_model):
        if rf.frame_basis is None:
            return results_y
        M = rf.vectorized_frame_basis
    elif isinstance(generator_model, LindbladModel):
        return rf.operator_out_of_frame_basis(results_y)
    else:
        if rf.frame_basis is None:
            return results_y
        M = rf.frame_basis
    if y0_ndim == 1:  # (T, n): one GEMM on the transposed block
        return _abi.zgemm(M, results_y.transpose(0, 1).contiguous()).transpose(0, 1).contiguous()
    return rf._left_multiply(M, results_y)


def solve_ode(rhs: Union[Callable, BaseGeneratorModel], t_span, y0, method="RK4", t_eval=None, **kwargs) -> OdeResult:
    """dy/dt = f(t, y).  ``rhs`` is a model or a callable returning device tensors."""
    if method not in ODE_METHODS:
        raise _unsupported(method, "solve_ode")
    if "max_dt" not in kwargs:
        raise QiskitError("fixed-step method RK4 requires max_dt.")
    y0 = asarray(y0)
    if not isinstance(rhs, BaseGeneratorModel):
        return RK4_solver(rhs, t_span, y0, t_eval=t_eval, **kwargs)

    _, solver_rhs, y0_fb, was_in_frame_basis = setup_generator_model_rhs_y0_in_frame_basis(rhs, y0)
    try:
        if _has_linear_generator(rhs):
            results = rk4_model_solve(rhs, t_span, y0_fb, t_eval=t_eval, **kwargs)
        elif is_lindblad_model_not_vectorized(rhs) and _abi.lindblad_supported(rhs.dim) and y0_fb.ndim in (2, 3):
            # non-vectorised Lindblad, dim <= 32: step loop on the device, density matrices resident on chip
            results = lindblad_rk4_model_solve(rhs, t_span, y0_fb, t_eval=t_eval, **kwargs)
        else:  # larger non-vectorised Lindblad systems: host-driven RK4 over the GEMM-based collection
            results = RK4_solver(solver_rhs, t_span, y0_fb, t_eval=t_eval, **kwargs)
        if not was_in_frame_basis:
            results.y = results_y_out_of_frame_basis(rhs, results.y, y0.ndim)
    finally:
        rhs.in_frame_basis = was_in_frame_basis
        wait_pending_copies()
    return results


def solve_lmde(generator: Union[Callable, BaseGeneratorModel], t_span, y0, method="RK4", t_eval=None, **kwargs) -> OdeResult:
    """dy/dt = G(t) y (solvers/solver_functions.py:220-373)."""
    if method in ODE_METHODS:
        if isinstance(generator, BaseGeneratorModel):
            rhs = generator
        else:
            def rhs(t, y):
                G = asarray(generator(t))
                y2 = y.reshape(-1, 1) if y.ndim == 1 else y
                out = _abi.zgemm(G.contiguous(), y2.contiguous())
                return out.reshape(-1) if y.ndim == 1 else out
        return solve_ode(rhs, t_span, y0, method=method, t_eval=t_eval, **kwargs)
    if method not in LMDE_METHODS:
        raise _unsupported(method, "solve_lmde")
    if is_lindblad_model_not_vectorized(generator):
        raise QiskitError("LMDE-specific methods with LindbladModel requires setting a vectorized=True.")
    if "max_dt" not in kwargs:
        raise QiskitError(f"fixed-step method {method} requires max_dt.")
    y0 = asarray(y0)
    if method in PARALLEL_METHODS and not isinstance(generator, BaseGeneratorModel):
        raise QiskitError(f"Method {method} needs a model generator (GeneratorModel, HamiltonianModel or a vectorized "
                          "LindbladModel): the step propagators are built from its operators on the device.")
    if not isinstance(generator, BaseGeneratorModel):
        return scipy_expm_solver(generator, t_span, y0, t_eval=t_eval, **kwargs)

    _, _, y0_fb, was_in_frame_basis = setup_generator_model_rhs_y0_in_frame_basis(generator, y0)
    try:
        if method in PARALLEL_METHODS:
            if PARALLEL_METHODS[method] == "RK4" and "magnus_order" in kwargs:
                raise QiskitError("magnus_order is an option of the exponential solvers only.")
            results = parallel_model_solve(generator, t_span, y0_fb, t_eval=t_eval, kind=PARALLEL_METHODS[method], **kwargs)
        else:
            results = expm_model_solve(generator, t_span, y0_fb, t_eval=t_eval, **kwargs)
        if not was_in_frame_basis:
            results.y = results_y_out_of_frame_basis(generator, results.y, y0.ndim)
    finally:
        generator.in_frame_basis = was_in_frame_basis
        wait_pending_copies()
    return results
