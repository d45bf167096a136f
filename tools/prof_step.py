#!/usr/bin/env python3
"""Run N fused loop steps of one bench workload for a profiler.

    python tools/prof_step.py <workload> <steps> [nograph] [spin=K]

`nograph` runs the steps in stream mode (qd_use_graphs(0)): ncu cannot open the kernel nodes of the production step
graph, which contains a conditional WHILE node.  `spin=K` runs K graph-mode steps first (developed winds / currents)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    name, nsteps = sys.argv[1], int(sys.argv[2])
    nograph = "nograph" in sys.argv[3:]
    spin = next((int(a.split("=")[1]) for a in sys.argv[3:] if a.startswith("spin=")), 4)
    import torch
    from qingdai_b200.simulation import Simulation
    spec = bench.workload(name)
    members = spec["members_total"] or 1
    ins = [bench.member_inputs(spec, m) for m in range(members)]
    sim = Simulation(spec["nlat"], spec["nlon"], [t for t, _ in ins], [p for _, p in ins], dt=spec["dt"], batch=members,
                     loop_with_albedo=spec["with_albedo"], device="cuda:0")
    sim.step(spin)
    torch.cuda.synchronize()
    if nograph:
        sim.engine.use_graphs(0)
    for _ in range(nsteps):
        sim.step(1)
    torch.cuda.synchronize()
    print("prof_step done:", name, nsteps, "n_sub", sim.engine.last_nsub()[:4], "launches", sim.engine.launches())


if __name__ == "__main__":
    main()
