"""Host logic, CPU only: the hierarchy connectivity the library derives at vrt_set_hierarchy (vrt_conn_*, vrt_amr.cu)
against the oracle's restatement of Rectangle::CalculateConnectivitySame / FromFiner (Rectangle.cpp:671-864), on every
hierarchy the reference produced in the AMR fixtures (3 levels, regridded, up to 6 adjacent finest patches)."""
import numpy as np
import pytest

from common import load_golden, species_from, meta
from oracle.port import MeshOracle, hierarchy_from_dump
from veritas_b200.solver import connectivity


@pytest.mark.parametrize("name", ["amr3_48x32_regrid", "amr2_tail_64x48_steps", "amr2_64x32_stages"])
def test_connectivity_matches_oracle(name):
    d = load_golden(name)
    mt = meta(d)
    maxd = mt["Lfinest"] - 1
    seen = set()
    n_same = n_coarse = 0
    for n in range(0, mt["steps"] + 1):
        H = hierarchy_from_dump(d, f"step{n}")
        O = MeshOracle(mt["nx"] * 2 ** maxd, mt["dx"], species_from(d), H, r=2, max_depth=maxd, poisson=False)
        for s in range(2):
            key = tuple(tuple(sorted((k, v) for k, v in p.items() if k != "key")) for p in H[s])
            if key in seen:
                continue
            seen.add(key)
            conn = connectivity(H[s], r=2, max_depth=maxd)
            for p in range(len(H[s])):
                for side in range(4):
                    nb, same = O.strips(s, p, side)
                    assert conn[p]["nb"][side] == nb, (n, s, p, side)
                    assert conn[p]["same"][side] == same, (n, s, p, side)
                    n_same += sum(1 for a, b in zip(nb, same) if a >= 0 and b)
                    n_coarse += sum(1 for a, b in zip(nb, same) if a >= 0 and not b)
                assert np.array_equal(conn[p]["flags"], O.flags(s, p)), (n, s, p)
    assert n_coarse > 0
    if name == "amr3_48x32_regrid":
        assert n_same > 0 and len(seen) >= 4


def test_connectivity_rejects_bad_hierarchies():
    from veritas_b200 import VrtError
    base = dict(depth=1, x_pos=0, p_pos=0, n_x=16, n_p=8, up=1, down=1, left=1, right=1)
    with pytest.raises(VrtError):
        connectivity([base, dict(base, depth=0, x_pos=3, p_pos=2, n_x=4, n_p=4)], r=2, max_depth=1)   # unaligned fine patch
    with pytest.raises(VrtError):
        connectivity([dict(base, n_x=5)], r=2, max_depth=1)
