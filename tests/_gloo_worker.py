"""Worker for test_gloo_world_size_2_gather (launched by torch.distributed.run, gloo backend)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_dynamics_b200 import distributed as D  # noqa: E402

dist.init_process_group("gloo")
rank, w = D.world()
assert w == 2
torch.manual_seed(0)
n, B, T = 5, 7, 3  # ragged: 4 + 3 columns
full = torch.randn(T, n, B, dtype=torch.float64) + 1j * torch.randn(T, n, B, dtype=torch.float64)
lo, hi = D.shard_bounds(B)
local = D.shard_columns(full)
assert local.shape == (T, n, hi - lo) and local.is_contiguous()
# pretend each rank evolved its block (a rank-independent linear map), then gather once
evolved = local * (2.0 + 1j)
gathered = D.all_gather_columns(evolved, B)
assert gathered.shape == full.shape
assert torch.equal(gathered, full * (2.0 + 1j))
obs = D.all_gather_columns(evolved.abs().sum(dim=(0, 1)), B)  # real per-column observable
assert torch.allclose(obs, (full * (2.0 + 1j)).abs().sum(dim=(0, 1)))
# even split (the bench's case): one all_gather_into_tensor, rank-major blocks land in column order
full8 = torch.randn(T, n, 8, dtype=torch.float64) + 1j * torch.randn(T, n, 8, dtype=torch.float64)
assert torch.equal(D.all_gather_columns(D.shard_columns(full8), 8), full8)
assert torch.equal(D.all_gather_columns(D.shard_columns(full8[0].real.contiguous()), 8), full8[0].real)
assert torch.equal(D.all_gather_columns(D.shard_columns(full8[0, 0].real.contiguous()), 8), full8[0, 0].real)

# a list of simulations split over the ranks (cfg5: parameter sweep sharded over the GPUs): every rank solves its block
# through the Solver protocol (a stand-in here: the product has no CPU path), one gather of final states / observables
class _Result:
    def __init__(self, t, y):
        self.t, self.y = t, y


class _StubSolver:
    calls = 0

    def solve(self, t_span, y0, signals, **kwargs):
        _StubSolver.calls += 1
        y0s = y0 if isinstance(y0, list) else [y0] * len(signals)
        return [_Result(torch.tensor(t_span, dtype=torch.float64), torch.stack([y, y * complex(s, -s)])) for y, s in zip(y0s, signals)]


class _StubMeasurement:
    def probabilities(self, t, Y):
        assert t == 2.0
        return torch.stack([Y.abs().sum(dim=0), (Y.real ** 2).sum(dim=0)])


nsim = 5  # ragged: 3 + 2
sigs = [float(k + 1) for k in range(nsim)]
yv = torch.arange(1, n + 1, dtype=torch.float64) + 0j
local, gathered = D.solver_solve_sharded(_StubSolver(), [0.0, 2.0], yv, sigs)
assert len(local) == len(D.shard_list(sigs)) == (3 if rank == 0 else 2)
expect = torch.stack([yv * complex(s, -s) for s in sigs], dim=-1)
assert gathered.shape == (n, nsim) and torch.equal(gathered, expect)
y0_list = [yv * (k + 1) for k in range(nsim)]
_, probs = D.solver_solve_sharded(_StubSolver(), [0.0, 2.0], y0_list, sigs, measurement=_StubMeasurement())
expect2 = torch.stack([yv * (k + 1) * complex(s, -s) for k, s in enumerate(sigs)], dim=-1)
assert probs.shape == (2, nsim) and torch.allclose(probs, torch.stack([expect2.abs().sum(dim=0), (expect2.real ** 2).sum(dim=0)]))
# fewer simulations than ranks: the idle rank still joins the gather
_, one = D.solver_solve_sharded(_StubSolver(), [0.0, 2.0], yv, [3.0])
assert one.shape == (n, 1) and torch.equal(one[:, 0], yv * complex(3.0, -3.0))
_, none = D.solver_solve_sharded(_StubSolver(), [0.0, 2.0], yv, sigs, gather=False)
assert none is None
dist.barrier()
print(f"GLOO_OK rank={rank}", flush=True)
dist.destroy_process_group()
