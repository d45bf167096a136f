"""Latitude bands on real GPUs (needs >= 2 devices; skipped on a single-GPU box): two ranks, one B200 each,
exchange halos / partial sums / median candidates through CUDA-IPC mapped buffers with peer-to-peer stores.
The assembled fields must match a single-GPU run of the same library (which the other GPU tests pin to the
reference) -- including the warp-streaming del^4 kernel on band segments and the whole-step CUDA graph."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("u", "v", "h", "ts", "q", "cloud", "hice", "uo", "vo", "eta", "sst", "precip", "albedo", "wland")


def _run(rank, world, port, out, nlat, nlon, dt, nsteps):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from qingdai_b200.params import QDParams
    from qingdai_b200.simulation import Simulation
    from qingdai_b200.synthetic import make_topography
    torch.cuda.set_device(rank)
    band = None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        band = (rank, world, 16)
    topo = make_topography(nlat, nlon, seed=42, land_frac=0.40)
    p = QDParams(energy_w=1.0, orog_enabled=True, cloud_couple=True)
    sim = Simulation(nlat, nlon, topo, p, dt=dt, loop_with_albedo=True, device=f"cuda:{rank}", band=band)
    for _ in range(nsteps):
        sim.step(1)
    full = {k: sim.engine.gather_rows(k) for k in FIELDS}
    err = sim.engine.band_info()[3]
    if rank == 0:
        np.savez(out, err=np.array(err), **full)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("nlat,nlon,dt,nsteps", [(181, 360, 300, 8), (721, 1440, 75, 4)])
def test_two_gpu_bands_match_one_gpu(tmp_path, nlat, nlon, dt, nsteps):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    mp.spawn(_run, args=(1, 0, one, nlat, nlon, dt, nsteps), nprocs=1, join=True)
    port = 29300 + (os.getpid() % 1500)
    mp.spawn(_run, args=(2, port, two, nlat, nlon, dt, nsteps), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    assert int(b["err"]) == 0
    for k in FIELDS:
        scale = max(float(np.max(np.abs(a[k]))), 1e-300)
        assert float(np.max(np.abs(a[k] - b[k]))) / scale < 1e-10, k
