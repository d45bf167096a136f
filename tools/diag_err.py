"""Where does a field deviate from the oracle?  python tools/diag_err.py nlat nlon dt spin [fused]"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import qdcheck
from qdcheck import model, QDParams, oracle_state_from_engine, reference_topography
from qingdai_b200._binding import default_library
from qingdai_b200.simulation import Simulation
nlat, nlon, dt, spin = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
lib = default_library()
p = QDParams(energy_w=1.0, orog_enabled=True, cloud_couple=True)
topo = reference_topography(nlat, nlon)
sim = Simulation(nlat, nlon, topo, p, dt=dt, lib=lib, loop_with_albedo=True)
eng = sim.engine
if "fused" in sys.argv:
    eng._chk(eng.lib.qd_set_ocean_fused(eng.ctx, 1), "fused")
g = model.make_grid(nlat, nlon)
Ts = 250.0 + 48.0 * (np.cos(np.deg2rad(g.lat)) ** 2)[:, None] * np.ones((nlat, nlon))
eng.set("ts", Ts); eng.set("sst", np.where(topo["land_mask"] == 0, Ts, 288.0))
eng.set("hice", np.where((topo["land_mask"] == 0) & (Ts < 268.0), 0.02, 0.0))
sim.step(spin)
st, oc = oracle_state_from_engine(eng, g, p, topo)
pre_cloud = st.cloud.copy()
t = sim.t
sim.step(1)
out = model.loop_step(st, oc, g, p, t=t, dt=dt, with_albedo_arg=True)
for mine, ref in (("cloud", st.cloud), ("precip", out.precip), ("u", st.u), ("v", st.v), ("q", st.q), ("ts", st.T_s), ("h", st.h)):
    got = eng.get(mine)
    d = np.abs(got - ref) / max(float(np.max(np.abs(ref))), 1e-300)
    j, i = np.unravel_index(np.argmax(d), d.shape)
    rows = d.max(axis=1)
    top = np.argsort(rows)[-6:][::-1]
    print(mine, "max", float(d.max()), "at", (int(j), int(i)), "lat", float(g.lat[j]), "val", float(ref[j, i]), "got-ref", float(got[j, i] - ref[j, i]),
          "top rows", [(int(r), float(rows[r])) for r in top], flush=True)
