"""GPU parity tests: the CUDA library (through its C ABI) against the oracle and the reference's golden files.

Run with `pytest -m gpu` on a B200 box.  The C ABI is the one declared in include/ipb200.h.
"""
import numpy as np
import pytest

import ipcases as ic
from oracle_lib import oracle, product
from parity import assert_same_call
from test_oracle_golden import (check_gdswzd_c, check_scalar, check_station_scalar, check_station_vector, check_vector)

pytestmark = pytest.mark.gpu
KINDS = ["4", "d", "8"]


# ---- the reference's own ctest cases, through the CUDA library --------------------------------------
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("grib", ["grib2", "grib1"])
@pytest.mark.parametrize("grid,ip", ic.SCALAR_CASES)
def test_scalar_golden_gpu(kind, grib, grid, ip):
    # device binary64 libm differs from glibc by <= 2 ulp, which occasionally flips the fp32 cast of the test
    check_scalar(product(kind), kind, grib, grid, ip, exact_frac=0.999)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("grib", ["grib2", "grib1"])
@pytest.mark.parametrize("grid,ip", ic.VECTOR_CASES)
def test_vector_golden_gpu(kind, grib, grid, ip):
    check_vector(product(kind), kind, grib, grid, ip, exact_frac=0.999)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ip", [0, 1, 2, 3, 6])
def test_station_points_gpu(kind, ip):
    a = check_station_scalar(product(kind), kind, "grib2", ip)
    b = check_station_scalar(oracle(kind), kind, "grib2", ip)
    assert_same_call(kind, a, b, label=f"station s ip{ip}")
    a = check_station_vector(product(kind), kind, "grib2", ip)
    b = check_station_vector(oracle(kind), kind, "grib2", ip)
    assert_same_call(kind, a, b, vector=True, label=f"station v ip{ip}")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("grib", ["grib2", "grib1"])
def test_gdswzd_c_entry_gpu(kind, grib):
    a = check_gdswzd_c(product(kind), grib)
    b = check_gdswzd_c(oracle(kind), grib)
    assert a["nret"] == b["nret"]
    tol = 2e-5 if kind == "4" else 1e-11
    for k in ("rlon", "rlat", "crot", "srot", "xlon", "xlat", "ylon", "ylat", "area"):
        assert np.allclose(a[k], b[k], rtol=tol, atol=tol), k
        if kind == "4":
            assert np.mean(a[k] == b[k]) > 0.999, (k, np.mean(a[k] == b[k]))


# ---- CUDA vs oracle, bit level, on cases the goldens do not reach (km>1, bitmaps, options) ----------
def _fields(mi, km, seed, kind, vector=False):
    rng = np.random.default_rng(seed)
    R = np.float32 if kind == "4" else np.float64
    base = ic.inputs()["u" if vector else "snoalb"][:mi]
    if base.size < mi:
        base = np.resize(base, mi)
    g = np.stack([base * (1 + 0.1 * k) + rng.standard_normal(mi) for k in range(km)]).astype(R)
    v = np.stack([rng.standard_normal(mi) * 5 + k for k in range(km)]).astype(R)
    # land-mask like bitmap: 8x8 blobs, ~70 % true
    li = np.zeros((km, mi), np.uint8)
    im = 360
    jj, ii = np.divmod(np.arange(mi), im)
    for k in range(km):
        h = ((ii >> 3) * 73856093 ^ (jj >> 3) * 19349663 ^ (k * 83492791)) % 10
        li[k] = (h < 7).astype(np.uint8)
    return g, v, li


CASES = [
    # (method, ipopt overrides {index(1-based): value}, output grid, km)
    (0, {}, "218", 5), (0, {1: 75}, "218", 3), (0, {2: 3}, "212", 4), (0, {}, "3", 3), (0, {}, "127", 2), (0, {}, "203", 2),
    (0, {}, "205", 2), (0, {}, "8", 3),
    (1, {}, "8", 4), (1, {1: 1}, "218", 3), (1, {2: 90}, "212", 2), (1, {}, "3", 2),
    (2, {}, "127", 4), (2, {1: 4}, "218", 3), (2, {}, "203", 2), (2, {}, "3", 2),
    (3, {}, "203", 3), (3, {}, "218", 3), (3, {1: 1, 2: 1, 3: 1, 4: 40}, "212", 2), (3, {1: 2, 2: -2, 5: 30}, "218", 2),
    (3, {20: 3}, "212", 3), (3, {}, "3", 2),
    (3, {1: 2, 2: 1, 3: 1, 4: 1, 5: 100}, "218", 4), (3, {1: 1, 2: 1, 3: 1, 4: 100}, "212", 3),   # mp = 100: lo decided at the threshold
    (6, {}, "212", 3), (6, {1: 1, 2: 2, 3: 1, 4: 50}, "218", 2), (6, {}, "3", 2),
]


def _ipopt(ip, over):
    o = ic.ipopt_for(ip)
    for k, v in over.items():
        o[k - 1] = v
    return o


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ip,over,grid,km", CASES)
def test_scalar_vs_oracle(kind, ip, over, grid, km):
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)[grid]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    g, _, li = _fields(mi, km, 11 + ip, kind)
    ibi = [(k % 2) for k in range(km)]
    opt = _ipopt(ip, over)
    a = product(kind).ipolates(ip, opt, gin, gout, mi, mo, km, ibi, li, g)
    b = oracle(kind).ipolates(ip, opt, gin, gout, mi, mo, km, ibi, li, g)
    st = assert_same_call(kind, a, b, exact=(ip == 2), label=f"s ip{ip} {grid}")
    print(kind, ip, over, grid, st)


VCASES = [(0, {}, "218", 3), (0, {}, "3", 2), (0, {}, "212", 2), (0, {}, "203", 2), (0, {}, "205", 2), (1, {1: 1}, "8", 2),
          (1, {}, "218", 2), (2, {}, "127", 2), (2, {1: 3}, "212", 2), (3, {}, "212", 2), (3, {}, "3", 2), (6, {}, "218", 2),
          (6, {}, "203", 2)]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ip,over,grid,km", VCASES)
def test_vector_vs_oracle(kind, ip, over, grid, km):
    gin, gout = ic.g2_input(kind, True), ic.g2_templates(kind, True)[grid]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    u, v, li = _fields(mi, km, 23 + ip, kind, vector=True)
    ibi = [(k % 2) for k in range(km)]
    opt = _ipopt(ip, over)
    a = product(kind).ipolatev(ip, opt, gin, gout, mi, mo, km, ibi, li, u, v)
    b = oracle(kind).ipolatev(ip, opt, gin, gout, mi, mo, km, ibi, li, u, v)
    st = assert_same_call(kind, a, b, vector=True, exact=False, label=f"v ip{ip} {grid}")
    print(kind, ip, over, grid, st)


@pytest.mark.parametrize("kind", ["4", "d"])
def test_error_codes(kind):
    """iret semantics: 32 bad option, 3 output grid too small, 2 no overlap."""
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)["218"]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    g, _, li = _fields(mi, 1, 3, kind)
    for ip, opt, mo_ in ((0, [101] + [0] * 19, mo), (1, [0, 200] + [0] * 18, mo), (3, [17] + [-1] * 19, mo), (0, [0] * 20, mo - 5),
                         (6, [-3] + [-1] * 19, mo)):
        a = product(kind).ipolates(ip, opt, gin, gout, mi, mo_, 1, [0], li, g)
        b = oracle(kind).ipolates(ip, opt, gin, gout, mi, mo_, 1, [0], li, g)
        assert a["iret"] == b["iret"] and a["no"] == b["no"], (ip, opt, a["iret"], b["iret"], a["no"], b["no"])
        assert a["iret"] != 0
    # regional input that does not cover the output: iret=2
    gin2 = ic.g2_templates(kind)["218"]
    gout2 = ("grib2", 20, [6, 255, -1, 255, -1, 255, -1, 32, 32, -60000000, 145000000, 56, -60000000, 280000000, 47625000, 47625000, 128, 0])
    mi2, mo2 = ic.npts_of(gin2), ic.npts_of(gout2)
    g2 = np.ones((1, mi2))
    a = product(kind).ipolates(0, [0] * 20, gin2, gout2, mi2, mo2, 1, [0], np.ones((1, mi2), np.uint8), g2)
    b = oracle(kind).ipolates(0, [0] * 20, gin2, gout2, mi2, mo2, 1, [0], np.ones((1, mi2), np.uint8), g2)
    assert a["iret"] == b["iret"] == 2 and a["no"] == b["no"] == 0


def test_plan_cache_reuse():
    lib = product("4")
    import ctypes as C
    lib.lib.ipgpu_plan_clear()
    lib.lib.ipgpu_plan_count.restype = C.c_longlong
    gin, gout = ic.g2_input("4"), ic.g2_templates("4")["218"]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    g, _, li = _fields(mi, 2, 5, "4")
    r1 = lib.ipolates(0, [0] * 20, gin, gout, mi, mo, 2, [0, 0], li, g)
    n1 = lib.lib.ipgpu_plan_count()
    r2 = lib.ipolates(0, [0] * 20, gin, gout, mi, mo, 2, [0, 0], li, g)
    hit = C.c_int(0)
    lib.lib.ipgpu_last_timing(None, None, C.byref(hit))
    assert n1 == 1 and lib.lib.ipgpu_plan_count() == 1 and hit.value == 1
    assert np.array_equal(r1["go"], r2["go"]) and np.array_equal(r1["lo"], r2["lo"])


# ---- the point-staged kernels (csrc/ipb_gtiles.cuh): whole passes of 8 fields, bitmap patterns, tile shapes ----
# (method, ipopt overrides, output grid, km, bitmap pattern)
#   none: ibi = 0 everywhere; same: every field carries the same bitmap (the usable-tap pattern is remembered);
#   blocks: the bitmap changes every 8 fields (pattern recomputed per pass); each: a different bitmap per field
#   (per-field path); mixed: ibi alternates 0/1
GT_CASES = [(0, {}, "218", 19, "same"), (0, {}, "218", 24, "blocks"), (0, {1: 75}, "212", 17, "each"), (0, {}, "203", 16, "mixed"),
            (0, {}, "3", 16, "same"), (0, {}, "127", 9, "same"),
            (2, {}, "218", 19, "each"), (2, {}, "203", 13, "mixed"), (2, {}, "212", 16, "none"),
            (1, {}, "218", 19, "none"), (1, {1: 1}, "8", 17, "mixed"), (1, {}, "203", 16, "none"), (1, {}, "212", 24, "same"),
            (3, {}, "218", 19, "same"), (3, {}, "218", 24, "blocks"), (3, {}, "212", 17, "none"), (3, {}, "203", 16, "each"),
            (3, {}, "218", 18, "mixed"), (3, {1: 1, 2: 1, 3: 1, 4: 40}, "212", 16, "same"), (3, {}, "3", 16, "same"),
            (3, {1: 2, 2: 1, 3: 1, 4: 1, 5: 100}, "218", 17, "same"), (3, {1: 2, 2: 1, 3: 1, 4: 1, 5: 100}, "218", 16, "mixed")]


def _gt_case(kind, ip, over, grid, km, pat):
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)[grid]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    g, _, li = _fields(mi, km, 31 + ip, kind)
    ibi = [1] * km
    if pat == "none":
        ibi = [0] * km
    elif pat == "mixed":
        ibi = [(k % 2) for k in range(km)]
    elif pat == "same":
        li[:] = li[0]
    elif pat == "blocks":
        for k in range(km):
            li[k] = li[(k // 8) * 8]
    opt = _ipopt(ip, over)
    a = product(kind).ipolates(ip, opt, gin, gout, mi, mo, km, ibi, li, g)
    b = oracle(kind).ipolates(ip, opt, gin, gout, mi, mo, km, ibi, li, g)
    return assert_same_call(kind, a, b, exact=False, label=f"gt ip{ip} {grid} {pat}")


@pytest.mark.parametrize("kind", ["4", "d"])
@pytest.mark.parametrize("ip,over,grid,km,pat", GT_CASES)
def test_staged_kernels_vs_oracle(kind, ip, over, grid, km, pat):
    st = _gt_case(kind, ip, over, grid, km, pat)
    if ip in (0, 1, 2):   # bilinear, bicubic, neighbor: same operations in the same order as the reference
        assert st[0]["exact_frac"] == 1.0, st
    print(kind, ip, grid, pat, st)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_staged_kernels_every_tile_shape(mode):
    """The tile shape is chosen per plan; force each one (read once per process, hence the child process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, IPB_GMODE=str(mode))
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(here, "test_gpu_parity.py"), "-k",
                        "test_staged_kernels_vs_oracle"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
