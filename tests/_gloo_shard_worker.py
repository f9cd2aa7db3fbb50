"""Worker for test_sharding_world_size_2_gloo: exercises rehuel_b200.sharding on CPU (gloo)."""
import sys

import numpy as np
import torch
import torch.distributed as dist

from oracle.refsolver import RefSolver, make_options
from rehuel_b200 import ensembles, sharding

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
M = 37  # deliberately not divisible
lo, hi = sharding.shard_range(M, rank, world)
assert sharding.shard_range(M, 0, 1) == (0, M)
covered = torch.zeros(M)
covered[lo:hi] = 1
dist.all_reduce(covered)
assert torch.all(covered == 1), "shards must tile [0, M) exactly once"

port = RefSolver("port")
o = make_options(rel_tol=1e-6, abs_tol=1e-6)
P = ensembles.robertson_params(hi - lo, start=lo)  # each rank generates ITS slice only
r = port.solve("robertson", 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, P, o)
local = sharding.pack_stats(attempt=r.counters["attempt"], steps=r.counters["attempt"] - r.counters["reject_err"],
                            newton_iters=np.zeros(hi - lo), fun_evals=r.counters["fun_evals"], n_failed=(r.status != 0),
                            members=hi - lo)
tot = sharding.allreduce_stats(local)
if rank == 0:
    full = port.solve("robertson", 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, ensembles.robertson_params(M), o)
    assert tot["members"] == M
    assert tot["attempt"] == int(full.counters["attempt"].sum())
    assert tot["fun_evals"] == int(full.counters["fun_evals"].sum())
    assert tot["max_attempt"] == int(full.counters["attempt"].max())
    assert tot["n_failed"] == 0
    print("SHARD-OK")
dist.destroy_process_group()
