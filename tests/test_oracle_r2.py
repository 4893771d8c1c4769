"""CPU checks of the oracle on the cases the reference's golden files do not pin ("unpinned"):
the inverse projections used on the source side and the elliptical polar-stereographic branch.  Pinned here by
size-independent properties: forward o inverse round trip (tests/reg_tests/gdswzd/sorc/gdswzd_driver.f90:296-330)
and interpolation of an analytic function of latitude/longitude from every projection."""
import numpy as np
import pytest

import ipcases as ic
import r2cases as rc
from oracle_lib import oracle

GRIDS = ["3", "218", "212", "212e", "8", "127", "205", "203"]


@pytest.mark.parametrize("kind", ["4", "d"])
@pytest.mark.parametrize("grid", GRIDS)
def test_oracle_gdswzd_round_trip(kind, grid):
    desc = rc.source_grids(kind)[grid]
    npts = ic.npts_of(desc)
    O = oracle(kind)
    f = O.gdswzd(desc, 0, npts)
    assert f["nret"] == npts
    b = O.gdswzd(desc, -1, npts, rlon=f["rlon"], rlat=f["rlat"])
    # a grid point exactly at the pole is rejected by the inverse's ABS(RLAT) < 90 + TINYREAL test
    # (ip_polar_stereo_grid_mod.F90:409): at most that one point may come back as fill
    ok = b["xpts"] != -9999.0
    assert b["nret"] == int(ok.sum()) >= npts - 1
    assert (np.abs(f["rlat"][~ok]) == 90.0).all()
    assert np.abs(b["xpts"] - f["xpts"])[ok].max() <= 0.01 and np.abs(b["ypts"] - f["ypts"])[ok].max() <= 0.01
    for k in ("crot", "srot", "xlon", "xlat", "ylon", "ylat", "area"):
        assert np.isfinite(b[k]).all() and np.isfinite(f[k]).all()
        tol = 2e-3 if kind == "4" else 1e-8
        assert np.allclose(b[k][ok], f[k][ok], rtol=tol, atol=tol * max(1.0, float(np.abs(f[k]).max()))), k


@pytest.mark.parametrize("grid", ["218", "212", "212e", "8", "127", "205", "203"])
def test_oracle_interpolates_analytic_field_from_every_projection(grid):
    kind = "d"
    gin, gout = rc.source_grids(kind)[grid], rc.conus_latlon(kind)
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    O = oracle(kind)
    src = O.gdswzd(gin, 0, mi)
    fun = lambda lat, lon: np.sin(np.radians(lat)) * np.cos(np.radians(lon)) + 0.3 * np.sin(np.radians(2 * lon))
    gi = fun(src["rlat"], src["rlon"])[None]
    r = O.ipolates(0, [0] * 20, gin, gout, mi, mo, 1, [0], np.ones((1, mi), np.uint8), gi)
    assert r["iret"] == 0 and r["no"] == mo
    ok = r["lo"][0] == 1
    assert ok.mean() > 0.5
    err = np.abs(r["go"][0] - fun(r["rlat"], r["rlon"]))[ok]
    assert err.max() < 2e-3, err.max()
    n = O.ipolates(2, [0] * 20, gin, gout, mi, mo, 1, [0], np.ones((1, mi), np.uint8), gi)
    assert np.abs(n["go"][0] - fun(n["rlat"], n["rlon"]))[n["lo"][0] == 1].max() < 3e-2


def test_oracle_earth_shape_scales_projection():
    kind = "d"
    O = oracle(kind)
    base = ic.g2_templates(kind)["218"]
    radii = {}
    for code in range(0, 9):
        d = rc.shaped(base, code, kind)
        r = O.gdswzd(d, 0, ic.npts_of(d))
        assert r["nret"] == ic.npts_of(d), code
        radii[code] = r["rlat"][-1]
    assert radii[6] != radii[0] and radii[1] != radii[6]
    # an unknown code gives radius -9999 (earth_radius_mod.F90:88-90); only the rotated grids reject it (nret = 0)
    bad = O.gdswzd(rc.shaped(ic.g2_templates(kind)["205"], 9, kind), 0, ic.npts_of(ic.g2_templates(kind)["205"]))
    assert bad["nret"] == 0
