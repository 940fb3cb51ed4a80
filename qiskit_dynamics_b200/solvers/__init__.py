"""Solvers (mirror of qiskit_dynamics.solvers): Solver, solve_lmde, solve_ode and the fixed-step grid."""
from .solver_functions import solve_lmde, solve_ode, ODE_METHODS, LMDE_METHODS
from .solver_classes import Solver
from .fixed_step import (RK4_solver, scipy_expm_solver, get_fixed_step_sizes, merge_t_args, trim_t_results,
                         stage_time_grid, rk4_model_solve, expm_model_solve)

__all__ = ["Solver", "solve_lmde", "solve_ode", "RK4_solver", "scipy_expm_solver", "get_fixed_step_sizes",
           "merge_t_args", "trim_t_results", "stage_time_grid", "rk4_model_solve", "expm_model_solve",
           "ODE_METHODS", "LMDE_METHODS"]
