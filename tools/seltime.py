import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import qdcheck
from qingdai_b200._binding import default_library
lib = default_library()
for shape in ((1441, 2880), (181, 360)):
    eng = qdcheck.make_engine(lib, *shape)
    rng = np.random.default_rng(0)
    x = np.exp(rng.standard_normal(shape) * 0.3); x[rng.random(shape) < 0.4] = 0.0
    t = torch.from_numpy(x).to(eng.device).contiguous()
    out = np.zeros(1)
    for rep in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng._chk(eng.lib.qd_median_pos(eng.ctx, qdcheck._ptr_of(t), -1.0, qdcheck._ptr_of(out)), "m")
        e1.record(); torch.cuda.synchronize()
        print(shape, rep, out[0] == np.median(x[x > 0]), "%.1f us" % (e0.elapsed_time(e1) * 1e3))
        print("   stats", eng.median_stats()[3])
        if rep == 3:
            t.mul_(4096.0); x = x * 4096.0
