"""Latitude-band decomposition on CPU (BASELINE configs[4]): gloo ranks each run the host check
build of the kernels on their block of latitude rows; halos, partial sums and the median candidates travel
through shared-memory exchange buffers (the CUDA build maps the same buffers with CUDA IPC and writes them over
NVLink).  The assembled fields must match a single-process run of the whole grid: bit-exact where no global
sum is involved, <= 1e-12 otherwise (the area-weighted sums are combined per rank, so their last bits differ)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NLAT, NLON, NSTEPS, DT = 49, 72, 7, 600
FIELDS = ("u", "v", "h", "ts", "q", "cloud", "hice", "uo", "vo", "eta", "sst", "precip", "albedo", "wland", "ssnow")


def _sim(lib, band=None):
    from qingdai_b200.params import QDParams
    from qingdai_b200.simulation import Simulation
    from qingdai_b200.synthetic import make_topography
    topo = make_topography(NLAT, NLON, seed=42, land_frac=0.35)
    p = QDParams(energy_w=1.0, orog_enabled=True, cloud_couple=True, shapiro_every=3, spec_every=4)
    return Simulation(NLAT, NLON, topo, p, dt=DT, lib=lib, loop_with_albedo=True, band=band)


def _run(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hostcheck import library
    band = None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        band = (rank, world, 8)
    sim = _sim(library(), band)
    sim.step(NSTEPS)
    full = {k: sim.engine.gather_rows(k) for k in FIELDS}
    nsub = sim.engine.scalars()[0]
    err = sim.engine.band_info()[3]
    if rank == 0:
        np.savez(out, err=np.array(err), **full)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_latitude_bands_match_single_process(tmp_path, world):
    one, many = str(tmp_path / "one.npz"), str(tmp_path / f"w{world}.npz")
    _run(0, 1, 0, one)
    port = 29700 + (os.getpid() % 1500) + world
    mp.spawn(_run, args=(world, port, many), nprocs=world, join=True)
    a, b = np.load(one), np.load(many)
    assert int(b["err"]) == 0
    for k in FIELDS:
        scale = max(float(np.max(np.abs(a[k]))), 1e-300)
        err = float(np.max(np.abs(a[k] - b[k]))) / scale
        assert err < 1e-10, (k, err)
