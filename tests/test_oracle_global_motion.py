"""oracle_obmc_render_ref pinned against the compiled reference's schro_motion_render with global motion on
(the dispatcher then takes schro_motion_render_ref, schroedinger/schromotion.c:113-121).  CPU only."""
import numpy as np
import pytest

from tests import helpers

ref = helpers.load_ref()
oracle = helpers.load_oracle()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")


@pytest.mark.parametrize("prec", [0, 1, 2, 3])
@pytest.mark.parametrize("add", [1, 0])
def test_global_motion_render_matches_reference(prec, add):
    rng = np.random.default_rng(40 + prec)
    case, gm = helpers.global_motion_case(oracle, 176, 144, rng, prec=prec)
    want = helpers.ref_obmc_global(ref, case, add, gm)
    got = helpers.oracle_obmc_ref(oracle, case, add, gm)
    for k in range(3):
        for part, name in enumerate(("acc", "residual", "out")):
            if name == "out" and not add:
                continue
            assert np.array_equal(got[k][part], want[k][part]), (prec, add, k, name)
    assert ((case.mvs["flags"] >> 2) & 1).sum() > 50


def test_global_motion_other_geometries_and_weights():
    rng = np.random.default_rng(7)
    for kw in (dict(xbsep=8, ybsep=8, xblen=8, yblen=8), dict(xbsep=16, ybsep=16, xblen=24, yblen=24, weights=(3, 1, 2)),
               dict(xbsep=4, ybsep=4, xblen=8, yblen=8, num_refs=1), dict(weights=(1, 3, 2), span=300)):
        case, gm = helpers.global_motion_case(oracle, 100, 70, rng, **kw)
        if kw.get("num_refs") == 1:
            assert not ((case.mvs["flags"] & 3) >= 2).any()          # make_mv_field: one reference -> modes 0 / 1 only
        for add in (1, 0):
            want = helpers.ref_obmc_global(ref, case, add, gm)
            got = helpers.oracle_obmc_ref(oracle, case, add, gm)
            for k in range(3):
                for part in range(3 if add else 2):
                    assert np.array_equal(got[k][part], want[k][part]), (kw, add, k, part)
