import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import qdcheck
from qingdai_b200._binding import default_library
lib=default_library()
try:
    qdcheck.check_ocean_fused_one_substep(lib); print('one substep exact OK', flush=True)
except AssertionError as e:
    print('one substep FAIL', str(e)[:300], flush=True)
for shape,dt,b in (((401,800),120.0,1),((401,800),200.0,1),((401,800),60.0,1),((181,360),300.0,32)):
    try:
        print(shape,dt,b, qdcheck.check_large_grid_paths_agree(lib, shape=shape, nsteps=4, dt=dt, batch=b), flush=True)
    except AssertionError as e:
        print(shape,dt,b,'FAIL',str(e)[:300], flush=True)
