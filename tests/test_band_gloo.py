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
HALO = 8
FIELDS = ("u", "v", "h", "ts", "q", "cloud", "hice", "uo", "vo", "eta", "sst", "precip", "albedo", "wland", "ssnow")


def _sim(lib, band=None):
    from qingdai_b200.params import QDParams
    from qingdai_b200.simulation import Simulation
    from qingdai_b200.synthetic import make_topography
    topo = make_topography(NLAT, NLON, seed=42, land_frac=0.35)
    p = QDParams(energy_w=1.0, orog_enabled=True, cloud_couple=True, shapiro_every=3, spec_every=4)
    return Simulation(NLAT, NLON, topo, p, dt=DT, lib=lib, loop_with_albedo=True, band=band)


def _run(rank, world, port, out, shape=None, halo=None, dt=None):
    global NLAT, NLON, HALO, DT
    if dt is not None:
        DT = dt
    if shape is not None:
        NLAT, NLON = shape
    if halo is not None:
        HALO = halo
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hostcheck import library
    band = None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        band = (rank, world, HALO)
    sim = _sim(library(), band)
    sim.step(NSTEPS)
    full = {k: sim.engine.gather_rows(k) for k in FIELDS}
    nsub = sim.engine.scalars()[0]
    err = sim.engine.band_info()[3]
    if rank == 0:
        np.savez(out, err=np.array(err), nsub=np.array(float(sim.engine.last_nsub()[0])), **full)
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


def test_latitude_bands_two_substeps_per_exchange(tmp_path):
    """Halo of 16 rows: one exchange serves a GROUP of two ocean sub-steps (8 halo rows each; qd_api.cu:
    ocean_group_size); dt chosen so that n_sub > 1 and odd / even counts both occur over the run."""
    global NLAT, NLON, HALO, DT
    one, many = str(tmp_path / "one.npz"), str(tmp_path / "w2.npz")
    saved = (NLAT, NLON, HALO, DT)
    shape = (97, 96)
    os.environ["QD_OCEAN_GROUP"] = "2"                    # opt-in (the default is one sub-step per exchange); inherited by the spawned ranks
    try:
        _run(0, 1, 0, one, shape, 16, 1500)
        port = 29700 + (os.getpid() % 1500) + 7
        mp.spawn(_run, args=(2, port, many, shape, 16, 1500), nprocs=2, join=True)
    finally:
        NLAT, NLON, HALO, DT = saved
        os.environ.pop("QD_OCEAN_GROUP", None)
    a, b = np.load(one), np.load(many)
    assert int(b["err"]) == 0
    assert float(a["nsub"]) >= 2.0, a["nsub"]
    for k in FIELDS:
        scale = max(float(np.max(np.abs(a[k]))), 1e-300)
        err = float(np.max(np.abs(a[k] - b[k]))) / scale
        assert err < 1e-10, (k, err)


def test_latitude_bands_with_trimmed_pole_ranks(tmp_path):
    """Unequal shares: the two pole ranks hand rows to the rank in between (qd_api.cu: band_rows_of; the default of 32
    rows only applies to shares of >= 128 rows, so the override is used here)."""
    one, many = str(tmp_path / "one.npz"), str(tmp_path / "w3.npz")
    os.environ["QD_BAND_POLE_TRIM"] = "5"                # inherited by the spawned ranks
    try:
        _run(0, 1, 0, one)
        port = 29700 + (os.getpid() % 1500) + 11
        mp.spawn(_run, args=(3, port, many), nprocs=3, join=True)
    finally:
        os.environ.pop("QD_BAND_POLE_TRIM", None)
    a, b = np.load(one), np.load(many)
    assert int(b["err"]) == 0
    for k in FIELDS:
        scale = max(float(np.max(np.abs(a[k]))), 1e-300)
        err = float(np.max(np.abs(a[k] - b[k]))) / scale
        assert err < 1e-10, (k, err)
