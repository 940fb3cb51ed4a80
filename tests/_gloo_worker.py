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
dist.barrier()
print(f"GLOO_OK rank={rank}", flush=True)
dist.destroy_process_group()
