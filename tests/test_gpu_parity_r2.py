"""GPU parity, round 2: the rows of SURVEY.md section 8 that had no CUDA-vs-oracle comparison yet.

All calls enter the CUDA library through the C ABI of include/ipb200.h (ctypes) and are compared with the CPU oracle
on the same inputs: lo / ibo / iret / no / geometry bit-exact, values bit-exact where the operation order is the
reference's (everything except collapsed budget in the binary64 kinds: <= 1e-12).  No tie point is forgiven
(tests/parity.py).
"""
import numpy as np
import pytest

import ipcases as ic
import r2cases as rc
from oracle_lib import oracle, product
from parity import RECORDS, assert_same_call

pytestmark = pytest.mark.gpu
KINDS = ["4", "d", "8"]


def _opt(ip):
    return ic.ipopt_for(ip)


# ---- (a) the four *_single_field entry points (tests/interp_mod_grib2.F90:250-261 calls both forms) ---------------
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("grib", ["grib2", "grib1"])
@pytest.mark.parametrize("grid,ip", [("218", 0), ("212", 6), ("3", 0), ("8", 1), ("127", 2), ("203", 3)])
def test_ipolates_single_field(kind, grib, grid, ip):
    gin = ic.g2_input(kind) if grib == "grib2" else ic.g1_input()
    gout = ic.g2_templates(kind)[grid] if grib == "grib2" else ic.g1_templates()[grid]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    g, _, li = rc.fields_for(mi, 1, 41 + ip, kind, im=360)
    for ibi in (0, 1):
        a = product(kind).ipolates(ip, _opt(ip), gin, gout, mi, mo, 1, [ibi], li, g, single=True)
        b = oracle(kind).ipolates(ip, _opt(ip), gin, gout, mi, mo, 1, [ibi], li, g)
        st = assert_same_call(kind, a, b, exact=(ip != 3 or kind == "4"), label=f"single s {grib} {grid} ip{ip} ibi{ibi} _{kind}")
        # and the array form of the product gives the same bits
        c = product(kind).ipolates(ip, _opt(ip), gin, gout, mi, mo, 1, [ibi], li, g)
        assert np.array_equal(a["go"], c["go"]) and np.array_equal(a["lo"], c["lo"]) and np.array_equal(a["ibo"], c["ibo"])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("grib", ["grib2", "grib1"])
@pytest.mark.parametrize("grid,ip", [("218", 0), ("212", 6), ("3", 0), ("8", 1), ("127", 2), ("203", 0)])
def test_ipolatev_single_field(kind, grib, grid, ip):
    gin = ic.g2_input(kind, True) if grib == "grib2" else ic.g1_input(True)
    gout = ic.g2_templates(kind, True)[grid] if grib == "grib2" else ic.g1_templates()[grid]   # grib1 203: kgds(11) |= 256 inside
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    u, v, li = rc.fields_for(mi, 1, 51 + ip, kind, im=360)
    for ibi in (0, 1):
        a = product(kind).ipolatev(ip, _opt(ip), gin, gout, mi, mo, 1, [ibi], li, u, v, single=True)
        b = oracle(kind).ipolatev(ip, _opt(ip), gin, gout, mi, mo, 1, [ibi], li, u, v)
        assert_same_call(kind, a, b, vector=True, exact=True, label=f"single v {grib} {grid} ip{ip} ibi{ibi} _{kind}")


# ---- (d) ipolatev_grib1 with kgds(1) = 203: the wind-point switch of src/ipolatev.F90:602-609 ---------------------
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ip", [0, 2, 3])
def test_ipolatev_grib1_egrid_203(kind, ip):
    g1 = ic.g1_templates()
    km = 2
    # 203 as the output grid
    gin, gout = ic.g1_input(True), g1["203"]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    u, v, li = rc.fields_for(mi, km, 61 + ip, kind, im=360)
    kg_before = list(gout[1])
    a = product(kind).ipolatev(ip, _opt(ip), gin, gout, mi, mo, km, [0, 1], li, u, v)
    b = oracle(kind).ipolatev(ip, _opt(ip), gin, gout, mi, mo, km, [0, 1], li, u, v)
    assert_same_call(kind, a, b, vector=True, exact=True, label=f"v grib1 ->203 ip{ip} _{kind}")
    assert list(gout[1]) == kg_before
    # the switch matters: mass points (scalar call on the same kgds) lie elsewhere
    s = product(kind).ipolates(0, [0] * 20, ic.g1_input(), gout, ic.npts_of(ic.g1_input()), mo, 1, [0],
                               np.ones(ic.npts_of(ic.g1_input()), np.uint8), np.ones(ic.npts_of(ic.g1_input())))
    assert not np.array_equal(s["rlon"], a["rlon"])
    # 203 as the input grid (wind points on the source side), regional lat-lon output
    gin2, gout2 = g1["203"], rc.conus_latlon_g1()
    mi2, mo2 = ic.npts_of(gin2), ic.npts_of(gout2)
    u2, v2, li2 = rc.fields_for(mi2, km, 71 + ip, kind, im=669)
    a = product(kind).ipolatev(ip, _opt(ip), gin2, gout2, mi2, mo2, km, [1, 0], li2, u2, v2)
    b = oracle(kind).ipolatev(ip, _opt(ip), gin2, gout2, mi2, mo2, km, [1, 0], li2, u2, v2)
    assert_same_call(kind, a, b, vector=True, exact=True, label=f"v grib1 203-> ip{ip} _{kind}")


# ---- (b) every projection as the INPUT grid: its inverse on the source side ---------------------------------------
SRC = ["218", "212", "212e", "8", "127", "205", "203"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("src", SRC)
@pytest.mark.parametrize("ip", [0, 1, 2, 3, 6])
def test_scalar_from_every_projection(kind, src, ip):
    gin = rc.source_grids(kind)[src]
    gout = rc.conus_latlon(kind)
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    km = 3
    g, _, li = rc.fields_for(mi, km, 81 + ip, kind, im=rc.dims_of(gin)[0])
    ibi = [0, 1, 0]
    a = product(kind).ipolates(ip, _opt(ip), gin, gout, mi, mo, km, ibi, li, g)
    b = oracle(kind).ipolates(ip, _opt(ip), gin, gout, mi, mo, km, ibi, li, g)
    assert b["iret"] == 0 and b["lo"].any(), (src, ip, b["iret"])
    st = assert_same_call(kind, a, b, exact=(ip != 3 or kind == "4"), label=f"s {src}->latlon ip{ip} _{kind}")
    print(src, ip, kind, st)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("src,dst", [("218", "212"), ("212e", "218"), ("127", "203"), ("205", "8"), ("203", "212e"), ("8", "205")])
@pytest.mark.parametrize("ip", [0, 2])
def test_scalar_projection_to_projection(kind, src, dst, ip):
    G = rc.source_grids(kind)
    gin, gout = G[src], G[dst]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    km = 2
    g, _, li = rc.fields_for(mi, km, 91 + ip, kind, im=rc.dims_of(gin)[0])
    a = product(kind).ipolates(ip, _opt(ip), gin, gout, mi, mo, km, [1, 0], li, g)
    b = oracle(kind).ipolates(ip, _opt(ip), gin, gout, mi, mo, km, [1, 0], li, g)
    assert a["iret"] == b["iret"]
    if b["iret"] == 0:
        assert_same_call(kind, a, b, exact=True, label=f"s {src}->{dst} ip{ip} _{kind}")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("src", SRC)
@pytest.mark.parametrize("ip", [0, 2, 3])
def test_vector_from_every_projection(kind, src, ip):
    gin = rc.source_grids(kind, True)[src]
    gout = rc.conus_latlon(kind)
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    km = 2
    u, v, li = rc.fields_for(mi, km, 101 + ip, kind, im=rc.dims_of(gin)[0])
    a = product(kind).ipolatev(ip, _opt(ip), gin, gout, mi, mo, km, [0, 1], li, u, v)
    b = oracle(kind).ipolatev(ip, _opt(ip), gin, gout, mi, mo, km, [0, 1], li, u, v)
    assert b["iret"] == 0 and b["lo"].any(), (src, ip, b["iret"])
    assert_same_call(kind, a, b, vector=True, exact=True, label=f"v {src}->latlon ip{ip} _{kind}")


# ---- gdswzd: iopt 0 / +1 / -1, all optional outputs, every projection; forward o inverse round trip ---------------
GD_KEYS = ("xpts", "ypts", "rlon", "rlat", "crot", "srot", "xlon", "xlat", "ylon", "ylat", "area")


def _same_gdswzd(a, b, label, kind):
    assert a["nret"] == b["nret"], (label, a["nret"], b["nret"])
    for k in GD_KEYS:
        nd = int((a[k] != b[k]).sum())
        RECORDS.append(dict(case=label, kind=kind, name=k, n=int(a[k].size), lo_mismatch=0, out_of_tol=nd, worst=0,
                            exact_frac=1.0 - nd / a[k].size,
                            points=[dict(index=(int(i),), got=float(a[k][i]), ref=float(b[k][i]), lo_got=1, lo_ref=1)
                                    for i in np.flatnonzero(a[k] != b[k])[:64]]))
        assert nd == 0, (label, k, nd, a[k][a[k] != b[k]][:4], b[k][a[k] != b[k]][:4])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("grid", ["3", "218", "212", "212e", "8", "127", "205", "203"])
def test_gdswzd_all_options_every_projection(kind, grid):
    desc = rc.source_grids(kind)[grid]
    npts = ic.npts_of(desc)
    P, O = product(kind), oracle(kind)
    a0, b0 = P.gdswzd(desc, 0, npts), O.gdswzd(desc, 0, npts)
    assert b0["nret"] == npts
    _same_gdswzd(a0, b0, f"gdswzd iopt0 {grid} _{kind}", kind)
    # iopt = +1 from fractional grid coordinates (some outside the grid -> fill)
    rng = np.random.default_rng(5)
    im, jm = rc.dims_of(desc)
    xs = rng.uniform(-2.0, a0["xpts"].max() + 3.0, 4000)
    ys = rng.uniform(-2.0, a0["ypts"].max() + 3.0, 4000)
    a1, b1 = P.gdswzd(desc, 1, 4000, xpts=xs, ypts=ys), O.gdswzd(desc, 1, 4000, xpts=xs, ypts=ys)
    _same_gdswzd(a1, b1, f"gdswzd iopt+1 {grid} _{kind}", kind)
    # iopt = -1 from the grid's own lat/lon: the round trip of gdswzd_driver.f90:296-330 (|dx|, |dy| <= 0.01)
    am, bm = P.gdswzd(desc, -1, npts, rlon=a0["rlon"], rlat=a0["rlat"]), O.gdswzd(desc, -1, npts, rlon=b0["rlon"], rlat=b0["rlat"])
    _same_gdswzd(am, bm, f"gdswzd iopt-1 {grid} _{kind}", kind)
    ok = am["xpts"] != -9999.0       # a grid point exactly at the pole is rejected by the polar inverse (ABS(RLAT) < 90 + TINYREAL)
    assert am["nret"] == int(ok.sum()) >= npts - 1
    tol = 0.01
    assert np.abs(am["xpts"] - a0["xpts"])[ok].max() <= tol and np.abs(am["ypts"] - a0["ypts"])[ok].max() <= tol
    # iopt = -1 from arbitrary lat/lon (many outside the domain)
    lon = rng.uniform(0, 360, 4000)
    lat = rng.uniform(-90, 90, 4000)
    ar, br = P.gdswzd(desc, -1, 4000, rlon=lon, rlat=lat), O.gdswzd(desc, -1, 4000, rlon=lon, rlat=lat)
    _same_gdswzd(ar, br, f"gdswzd iopt-1 random {grid} _{kind}", kind)


@pytest.mark.parametrize("kind", ["4", "d"])
def test_gdswzd_grib1_every_projection(kind):
    P, O = product(kind), oracle(kind)
    for grid, desc in ic.g1_templates().items():
        npts = ic.npts_of(desc)
        a0, b0 = P.gdswzd(desc, 0, npts), O.gdswzd(desc, 0, npts)
        _same_gdswzd(a0, b0, f"gdswzd grib1 iopt0 {grid} _{kind}", kind)
        am, bm = P.gdswzd(desc, -1, npts, rlon=a0["rlon"], rlat=a0["rlat"]), O.gdswzd(desc, -1, npts, rlon=b0["rlon"], rlat=b0["rlat"])
        _same_gdswzd(am, bm, f"gdswzd grib1 iopt-1 {grid} _{kind}", kind)


# ---- (c) every GRIB2 earth-shape code through the product (earth_radius_mod.F90:40-95) -----------------------------
@pytest.mark.parametrize("kind", ["4", "d"])
@pytest.mark.parametrize("code", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_earth_shape_codes_through_product(kind, code):
    P, O = product(kind), oracle(kind)
    G = ic.g2_templates(kind)
    for grid in ("218", "212", "8"):            # Lambert / polar stereographic (elliptical for 2,3,4,5,7) / Mercator use the radius
        desc = rc.shaped(G[grid], code, kind)
        npts = ic.npts_of(desc)
        a, b = P.gdswzd(desc, 0, npts), O.gdswzd(desc, 0, npts)
        _same_gdswzd(a, b, f"earth shape {code} {grid} _{kind}", kind)
        if code != 9:
            assert b["nret"] == npts
    # and as the output grid of an interpolation
    gin, gout = ic.g2_input(kind), rc.shaped(G["212"], code, kind)
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    g, _, li = rc.fields_for(mi, 2, 7, kind, im=360)
    a = P.ipolates(0, [0] * 20, gin, gout, mi, mo, 2, [0, 1], li, g)
    b = O.ipolates(0, [0] * 20, gin, gout, mi, mo, 2, [0, 1], li, g)
    assert a["iret"] == b["iret"] and a["no"] == b["no"]
    if b["iret"] == 0:
        assert_same_call(kind, a, b, exact=True, label=f"earth shape {code} ->212 _{kind}")


# ---- (e) BASELINE.json configs 2-5 on the real grids, km = 8 ... 24 ---------------------------
def _s_fields(km, kind, second=False):
    return np.stack([rc.synth_s(k, kind, second) for k in range(km)])


def test_baseline_config2_bilinear_S_to_L():
    kind, km = "4", 12
    G = ic.baseline_grids(kind)
    mi, mo = ic.npts_of(G["S"]), ic.npts_of(G["L"])
    g = _s_fields(km, kind)
    rng = np.random.default_rng(2)
    g[3] = rng.standard_normal(mi).astype(np.float32)          # white noise field
    li = np.ones((km, mi), np.uint8)
    a = product(kind).ipolates(0, [0] * 20, G["S"], G["L"], mi, mo, km, [0] * km, li, g)
    b = oracle(kind).ipolates(0, [0] * 20, G["S"], G["L"], mi, mo, km, [0] * km, li, g)
    st = assert_same_call(kind, a, b, exact=True, label="config2 bilinear S->L _4 km12")
    assert st[0]["exact_frac"] == 1.0 and a["no"] == 1905141


def test_baseline_config3_budget_S_to_L_d_bitmap():
    kind, km = "d", 8
    G = ic.baseline_grids(kind)
    mi, mo = ic.npts_of(G["S"]), ic.npts_of(G["L"])
    g = _s_fields(km, kind)
    li = np.stack([rc.synth_mask(0) for _ in range(km)])
    li[5] = rc.synth_mask(5)
    opt = [-1] * 20
    a = product(kind).ipolates(3, opt, G["S"], G["L"], mi, mo, km, [1] * km, li, g)
    b = oracle(kind).ipolates(3, opt, G["S"], G["L"], mi, mo, km, [1] * km, li, g)
    st = assert_same_call(kind, a, b, exact=False, label="config3 budget S->L _d bitmap km8")   # 1e-12 (collapsed taps), lo/ibo exact
    assert st[0]["lo_mismatch"] == 0 and not a["lo"].all()


def test_baseline_config4_vector_S_to_P():
    kind, km = "4", 8
    G = ic.baseline_grids(kind)
    mi, mo = ic.npts_of(G["S"]), ic.npts_of(G["P"])
    u, v = _s_fields(km, kind), _s_fields(km, kind, True)
    li = np.ones((km, mi), np.uint8)
    a = product(kind).ipolatev(0, [0] * 20, G["S"], G["P"], mi, mo, km, [0] * km, li, u, v)
    b = oracle(kind).ipolatev(0, [0] * 20, G["S"], G["P"], mi, mo, km, [0] * km, li, u, v)
    st = assert_same_call(kind, a, b, vector=True, exact=True, label="config4 vector bilinear S->P _4 km8")
    assert st[0]["exact_frac"] == 1.0 and st[1]["exact_frac"] == 1.0


@pytest.mark.parametrize("ip", [2, 1])
def test_baseline_config5_neighbor_bicubic_S_to_E(ip):
    kind, km = "4", 24 if ip == 2 else 16
    G = ic.baseline_grids(kind)
    mi, mo = ic.npts_of(G["S"]), ic.npts_of(G["E"])
    g = _s_fields(km, kind)
    li = np.ones((km, mi), np.uint8)
    a = product(kind).ipolates(ip, [0] * 20, G["S"], G["E"], mi, mo, km, [0] * km, li, g)
    b = oracle(kind).ipolates(ip, [0] * 20, G["S"], G["E"], mi, mo, km, [0] * km, li, g)
    st = assert_same_call(kind, a, b, exact=True, label=f"config5 ip{ip} S->E _4 km{km}")
    assert st[0]["exact_frac"] == 1.0
