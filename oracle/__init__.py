"""CPU oracle for the Qingdai per-timestep loop.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy restatement of the reference
algorithm (mountain/qingdai, ``pygcm/*`` and the loop body of
``scripts/run_simulation.py``).  It exists so that the CUDA path can be checked
against something that runs anywhere; it is never imported by ``qingdai_b200``.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.

Pinning: every function here is validated in this repository's CPU test-suite
against golden vectors produced by *importing and running the reference itself*
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).  The third-party
arithmetic the reference leans on (scipy.ndimage 1.18.1 ``map_coordinates``,
``gaussian_filter``, ``convolve``) is restated here in plain NumPy and checked
bit-for-bit against SciPy where SciPy is installed.
"""
