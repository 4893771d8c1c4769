"""CUDA grid expanders (ipxwafs, ipxwafs2, ipxwafs3, ipxetas) against the oracle, bit for bit.

Reference cases: /root/reference/tests/test_ipxwafs.F90:36-96; the E-grid cases follow src/ipxetas.F90:90-204."""
import numpy as np
import pytest

import xcases as xc
from oracle_lib import oracle, product

pytestmark = pytest.mark.gpu
KINDS = ["4", "d", "8"]


def same(a, b, keys):
    for q in keys:
        x, y = a[q], b[q]
        if isinstance(x, np.ndarray):
            assert x.dtype == y.dtype and np.array_equal(x.view(np.uint8), y.view(np.uint8)), q
        else:
            assert x == y, q


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_ipxwafs_known_answers_and_oracle(kind, variant):
    P, O = product(kind), oracle(kind)
    r, back = xc.wafs_round_trip(P, variant)
    ro, backo = xc.wafs_round_trip(O, variant)
    ref = xc.WAFS_REF3 if variant == 3 else xc.WAFS_REF12
    assert r["iret"] == 0 and back["iret"] == 0
    assert np.abs(r["data_full"][0, 2599:2609] - np.array(ref)).max() < 1e-6
    assert np.abs(back["data_thin"][0] - xc.wafs_thin_field(P.R)).max() < 1e-6
    keys = ["opt_pts", "tmpl_thin", "tmpl_full", "data_thin", "data_full", "iret"] + (["ib_thin", "ib_full", "bitmap_thin", "bitmap_full"] if variant > 1 else [])
    same(r, ro, keys)
    same(back, backo, keys)


@pytest.mark.parametrize("kind", ["4", "d"])
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_ipxwafs_many_fields_with_bitmaps(kind, variant):
    P, O = product(kind), oracle(kind)
    km = 7
    rng = np.random.default_rng(100 + variant)
    fields = (280 + 20 * rng.standard_normal((km, xc.NTHIN))).astype(P.R)
    bt = (rng.random((km, xc.NTHIN)) > 0.15).astype(np.uint8)
    ib = [1, 0, 1, 1, 0, 1, 1]
    bt[3] = 1                                     # a field that says "bitmap" and has no hole keeps ib_full = 0
    r, back = xc.wafs_round_trip(P, variant, km=km, fields=fields, ib_thin=ib, bitmap_thin=bt)
    ro, backo = xc.wafs_round_trip(O, variant, km=km, fields=fields, ib_thin=ib, bitmap_thin=bt)
    keys = ["opt_pts", "tmpl_thin", "tmpl_full", "data_thin", "data_full", "iret"] + (["ib_thin", "ib_full", "bitmap_thin", "bitmap_full"] if variant > 1 else [])
    same(r, ro, keys)
    same(back, backo, keys)
    if variant > 1:
        assert list(r["ib_full"]) == [1, 0, 1, 0, 0, 1, 1]


@pytest.mark.parametrize("kind", ["4", "d"])
def test_ipxwafs_south_to_north_and_errors(kind):
    P, O = product(kind), oracle(kind)
    tt = [0] * 19
    tt[8], tt[17], tt[18] = 73, 1250000, 64
    thin = xc.wafs_thin_field(P.R)[None]
    for lib_pair in [(P, O)]:
        a = lib_pair[0].ipxwafs(1, 1, 1, xc.NPWAFS[::-1], tt, thin, [0] * 19, np.zeros((1, xc.NFULL)))
        b = lib_pair[1].ipxwafs(1, 1, 1, xc.NPWAFS[::-1], tt, thin, [0] * 19, np.zeros((1, xc.NFULL)))
        same(a, b, ["opt_pts", "tmpl_thin", "tmpl_full", "data_full", "iret"])
        assert a["iret"] == 0
    z = [0] * 19
    for lib in (P, O):
        assert lib.ipxwafs(1, 1, 1, xc.NPWAFS, z, thin, z, np.zeros((1, xc.NFULL)))["iret"] == 1
        assert lib.ipxwafs(2, -1, 1, xc.NPWAFS, z, thin, z, np.zeros((1, xc.NFULL)))["iret"] == 1
        bad = list(xc.NPWAFS)
        bad[40] += 1
        assert lib.ipxwafs(3, 1, 1, bad, tt, thin, z, np.zeros((1, xc.NFULL)))["iret"] == 1
    # idir = 0 only checks sizes and leaves every array alone
    a = P.ipxwafs(1, 0, 1, xc.NPWAFS, tt, thin, z, np.full((1, xc.NFULL), 5.0))
    b = O.ipxwafs(1, 0, 1, xc.NPWAFS, tt, thin, z, np.full((1, xc.NFULL), 5.0))
    same(a, b, ["opt_pts", "tmpl_thin", "tmpl_full", "data_thin", "data_full", "iret"])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("scan,idir", [(68, 0), (72, 0), (68, 1), (64, -1), (64, -2)])
def test_ipxetas_matches_oracle(kind, scan, idir):
    P, O = product(kind), oracle(kind)
    tin, npi, npo = xc.egrid_case(scan, idir)
    rng = np.random.default_rng(scan * 7 + idir)
    data = 10 + rng.standard_normal(npi)
    for holes in (False, True):
        bm = (rng.random(npi) > 0.1).astype(np.uint8) if holes else np.ones(npi, np.uint8)
        a = P.ipxetas(idir, 1, tin, bm, data, npo)
        b = O.ipxetas(idir, 1, tin, bm, data, npo)
        assert a["iret"] == 0
        same(a, b, ["igdtnumo", "tmpl_out", "bitmap_out", "data_out", "iret"])


def test_ipxetas_error_codes():
    P, O = product("4"), oracle("4")
    tin, npi, npo = xc.egrid_case(68, 0)
    one, z = np.ones(npi, np.uint8), np.zeros(npi)
    for lib in (P, O):
        assert lib.ipxetas(0, 1, tin, one, z, npo + 1)["iret"] == 3
        assert lib.ipxetas(0, 0, tin, one, z, npo)["iret"] == 1
        assert lib.ipxetas(-1, 1, tin, one, z, npo)["iret"] == 2
    a, b = P.ipxetas(-1, 1, tin, one, z, npo), O.ipxetas(-1, 1, tin, one, z, npo)
    same(a, b, ["igdtnumo", "tmpl_out", "iret"])
