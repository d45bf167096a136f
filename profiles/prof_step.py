"""Small driver for ncu captures: N fused loop steps of a named bench workload (no timing, no CPU arm)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from qingdai_b200.simulation import Simulation  # noqa: E402
from qingdai_b200.synthetic import make_topography  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ensemble64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec = bench.workload(name)
members = spec["members_total"] or spec["members_per_gpu"]
topos = [make_topography(spec["nlat"], spec["nlon"], seed=42 + m, land_frac=0.40) for m in range(members)]
sim = Simulation(spec["nlat"], spec["nlon"], topos, spec["params"], dt=spec["dt"], batch=members, loop_with_albedo=True, device="cuda:0")
if len(sys.argv) > 3 and sys.argv[3] == "nograph":
    sim.engine.use_graphs(False)
for _ in range(steps):
    sim.step(1)
sim.engine.sync()
print("done", name, steps, "launches", sim.engine.launches())
