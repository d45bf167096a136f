"""N>1 path on CPU: two gloo ranks each run their share of a 4-member ensemble (host check build of the
kernels); the gathered per-member diagnostics must equal a single-process run of all four members bit
for bit -- members are independent, there is no data-path collective."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NLAT, NLON, NMEM, NSTEPS, DT = 19, 36, 4, 3, 600


def _topo(m):
    from qingdai_b200.synthetic import make_topography
    return make_topography(NLAT, NLON, seed=42 + m, land_frac=0.35)


def _params(m):
    from qingdai_b200.params import QDParams
    return QDParams(energy_w=1.0, gh_factor_lw=0.55 + 0.01 * m, sigma4=0.02 + 0.002 * m)


def _run(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hostcheck import library
    from qingdai_b200.ensemble import EnsembleRunner
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    run = EnsembleRunner(NLAT, NLON, NMEM, _topo, _params, dt=DT, rank=rank, world=world, lib=library(), loop_with_albedo=True)
    run.step(NSTEPS)
    diags = run.gather_diagnostics()
    if rank == 0:
        np.save(out, np.array([[d[k] for k in sorted(d)] for d in diags]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_member_slice_partitions():
    from qingdai_b200.ensemble import member_slice
    for n, w in [(64, 1), (64, 8), (5, 2), (3, 4)]:
        got = [m for r in range(w) for m in member_slice(n, w, r)]
        assert got == list(range(n))


def test_two_gloo_ranks_match_single_process(tmp_path):
    one, two = str(tmp_path / "one.npy"), str(tmp_path / "two.npy")
    _run(0, 1, 0, one)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_run, args=(2, port, two), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    assert a.shape == (NMEM, 8)
    assert np.array_equal(a, b)
    assert len({tuple(r) for r in a}) == NMEM          # members really differ (seeds / parameter sweep)
