"""Ensemble fan-out (BASELINE configs[3]): independent members are split across ranks, one
process per GPU, with NO data-path collective.  torch.distributed is used only to gather the
per-member diagnostics at the end (and, in bench.py, for the timing barrier / max-over-ranks)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .params import QDParams
from .simulation import Simulation


def member_slice(n_members: int, world: int, rank: int) -> range:
    """Contiguous block of members owned by ``rank`` (the remainder goes to the lowest ranks)."""
    base, rem = divmod(int(n_members), int(world))
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class EnsembleRunner:
    """Runs the members of one rank as a single batched Simulation (leading dimension of every kernel)."""

    def __init__(self, nlat, nlon, n_members, topo_for: Callable[[int], dict], params_for: Callable[[int], QDParams],
                 dt=300, rank=0, world=1, device=None, lib=None, **sim_kw):
        self.rank, self.world = int(rank), int(world)
        self.members = list(member_slice(n_members, world, rank))
        self.n_members = int(n_members)
        if not self.members:
            self.sim = None
            return
        topos = [topo_for(m) for m in self.members]
        params = [params_for(m) for m in self.members]
        self.sim = Simulation(nlat, nlon, topos, params, dt=dt, batch=len(self.members), device=device, lib=lib, **sim_kw)

    def step(self, nsteps=1):
        if self.sim is not None:
            self.sim.step(nsteps)

    def local_diagnostics(self) -> Dict[int, dict]:
        if self.sim is None:
            return {}
        return {m: self.sim.diagnostics(member=b) for b, m in enumerate(self.members)}

    def gather_diagnostics(self) -> Optional[List[dict]]:
        """All members' diagnostics on rank 0 (None elsewhere); the only communication of an ensemble run."""
        local = self.local_diagnostics()
        if self.world == 1:
            return [local[m] for m in range(self.n_members)]
        import torch.distributed as dist
        gathered = [None] * self.world if self.rank == 0 else None
        dist.gather_object(local, gathered, dst=0)
        if self.rank != 0:
            return None
        merged = {}
        for d in gathered:
            merged.update(d)
        return [merged[m] for m in range(self.n_members)]
