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
    """Runs the members of one rank as batched Simulations (leading dimension of every kernel).

    ``streams`` > 1 splits the rank's members into that many groups, each with its own context, CUDA stream and step
    graph: members are independent, so the groups' graphs run concurrently and the dependency bubbles of one group's
    step (a chain of ~60 short kernels at 181x360, three grid-wide syncs per exact median) are filled by the others.
    MEASURED on B200 (profiles/README.md, round 2): it does not pay -- 8 members: 0.51 ms/step on one stream, 0.58 / 0.73 /
    0.83 on 2 / 4 / 8; 16 members: 0.78 vs 0.84 / 0.99; 64 members: 2.59 vs 2.54.  The cooperative median launches of
    the groups serialise and the fork / join adds more than the bubbles give back.  Default: one stream."""

    def __init__(self, nlat, nlon, n_members, topo_for: Callable[[int], dict], params_for: Callable[[int], QDParams],
                 dt=300, rank=0, world=1, device=None, lib=None, streams=None, **sim_kw):
        self.rank, self.world = int(rank), int(world)
        self.members = list(member_slice(n_members, world, rank))
        self.n_members = int(n_members)
        self.sims, self.groups, self._streams = [], [], []
        if not self.members:
            self.sim = None
            return
        m = len(self.members)
        if streams is None:
            streams = 1
        k = max(1, min(int(streams), m))
        host_only = lib is not None and getattr(lib, "host_emulation", False)
        for g in range(k):
            grp = self.members[g * m // k:(g + 1) * m // k]
            if not grp:
                continue
            topos = [topo_for(i) for i in grp]
            params = [params_for(i) for i in grp]
            if k > 1 and not host_only:
                import torch
                st = torch.cuda.Stream(device=device)
                with torch.cuda.stream(st):              # the engine binds the stream that is current at construction
                    sim = Simulation(nlat, nlon, topos, params, dt=dt, batch=len(grp), device=device, lib=lib, **sim_kw)
                self._streams.append(st)
            else:
                sim = Simulation(nlat, nlon, topos, params, dt=dt, batch=len(grp), device=device, lib=lib, **sim_kw)
            self.sims.append(sim)
            self.groups.append(grp)
        self.sim = self.sims[0]

    def step(self, nsteps=1):
        for sim in self.sims:
            sim.step(nsteps)

    def synchronize(self):
        for sim in self.sims:
            sim.engine.sync()

    def local_diagnostics(self) -> Dict[int, dict]:
        out = {}
        for sim, grp in zip(self.sims, self.groups):
            for b, m in enumerate(grp):
                out[m] = sim.diagnostics(member=b)
        return out

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
