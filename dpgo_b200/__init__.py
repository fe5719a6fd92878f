"""dpgo_b200: B200-native MM-PGO / AMM-PGO* / AMM-PGO# iteration.

The arithmetic lives in libmmpgo.so (hand-written sm_100a FP64 kernels behind
the C ABI in include/mmpgo.h); this package is the thin host mirror of the
reference's driver interface plus graph I/O.  There is no CPU fallback.
"""
from .graph import PoseGraph, read_g2o, write_g2o, grid3d, sphere_rings, city2d  # noqa: F401
from .lib import MmpgoError, load  # noqa: F401
from .pgo import DPGOHash, DPGOStar, Options, project_to_SOdn, run_dist_pgo  # noqa: F401
