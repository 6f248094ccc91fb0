"""Cases and helpers shared by tests/golden/make_golden.py (fixture generator) and the tests that read
the fixtures.  The fixtures freeze OUTPUTS OF THE ORACLE (oracle/kamr_oracle.c): the reference is pure Julia +
libp4est + MPI and cannot run in this image, and its own tests hold no vectors, so these are a
regression pin of the restatement, not reference outputs."""
import hashlib

import numpy as np

from kitamr_jl_b200 import abi
from kitamr_jl_b200.synth import cases

CASES = {
    "S0": lambda: cases.smoke_s0(),
    "amr2d_ragged": lambda: cases.amr_case(dim=2, trees=4, maxlevel=2, vtrees=8, vs_maxlevel=2, ragged=True),
    "amr3d_ragged": lambda: cases.amr_case(dim=3, trees=3, maxlevel=1, vtrees=4, vs_maxlevel=1, ragged=True, seed=4),
    # immersed boundary (update_solid_cell! / update_solid_neighbor!, cut velocity cells), 2-D circle and 3-D sphere
    "s2_ib_small": lambda: cases.cylinder_s2(trees=5, ps_maxlevel=4, box_level=2, vtrees=8, vs_maxlevel=2, ib=True),
    "s4_ib_small": lambda: cases.sphere_s4(trees=4, ps_maxlevel=2, vtrees=4, vs_maxlevel=1),
    # S1 riemann2d under CIP_Marching, S3 airfoil2d (InterpolatedOutflow), S5 x38-like (cold wall, wide velocity box)
    "s1_small": lambda: cases.riemann_s1(ps_level=1, band_level=2, trees=4, vtrees=8, vs_maxlevel=2),
    "s3_small": lambda: cases.airfoil_s3(ps_maxlevel=3, box_level=2, trees=(6, 8), vtrees=12),
    "s5_small": lambda: cases.x38like_s5(ps_maxlevel=2, trees=4, vtrees=10, vs_maxlevel=1),
    "cip2d": lambda: cases.amr_case(dim=2, trees=4, maxlevel=1, vtrees=8, vs_maxlevel=1, ragged=True, seed=8,
                                    marching=abi.MARCH_CIP),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample_idx(n, k=4096, seed=99):
    rng = np.random.default_rng(seed)
    return np.sort(rng.choice(n, size=min(k, n), replace=False))
