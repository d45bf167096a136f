#!/usr/bin/env python3
"""Latency of the latitude-band halo exchange alone (no compute between exchanges), N ranks under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/band_micro.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from qingdai_b200.simulation import Simulation
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    spec = bench.workload("hires")
    topo, p = bench.member_inputs(spec, 0)
    for halo in (16, 8):
        sim = Simulation(spec["nlat"], spec["nlon"], [topo], [p], dt=spec["dt"], batch=1, loop_with_albedo=True, device=f"cuda:{local}", band=(rank, world, halo))
        eng = sim.engine
        for nf in (1, 4, 8):
            dist.barrier(); torch.cuda.synchronize()
            ms = ctypes.c_float(0)
            eng._chk(eng.lib.qd_band_exchange_bench(eng.ctx, nf, 200, ctypes.byref(ms)), "qd_band_exchange_bench")
            t = torch.tensor([ms.value], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"world {world} halo {halo} fields {nf}: {t.item() / 200 * 1e3:.2f} us per exchange ({nf * halo * spec['nlon'] * 8 / 1e6:.2f} MB per direction)", flush=True)
        del sim, eng
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
