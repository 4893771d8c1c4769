"""CPU-side checks of the boundary and of the arithmetic contract (no GPU needed).

* the C-ABI library loads without a device and exports every symbol include/ipb200.h declares;
* the product's correctly rounded binary64 functions (csrc/ipb_crmath.h, host build) and the oracle's
  independent implementation (oracle/ip_oracle_libm.hpp) both equal libquadmath's binary128 result rounded
  once, on random and adversarial arguments;
* how far the choice of libm alone moves the reference's results (oracle built with plain glibc vs the
  correctly rounded oracle): this is the noise floor any cross-platform comparison of the _d kind has.
"""
import ctypes as C
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

import ipcases as ic
from oracle_lib import ROOT, oracle

pkg = importlib.import_module("nceplibs-ip_b200")


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "ipb200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ipolate[sv]|gdswzd|ipgpu|ipxwafs[23]?|ipxetas)_[a-z0-9_]+|gdswzd_[4d8])\s*\(", h)))


def test_library_exports_every_declared_symbol():
    pkg.build_library()
    lib = C.CDLL(pkg.PRODUCT_SO)  # must load without a GPU
    names = _declared_symbols()
    assert len(names) >= 3 * 12 + 10, names
    for kind in "4d8":
        for stem in ("ipolates_grib2", "ipolates_grib1", "ipolates_grib2_single_field", "ipolates_grib1_single_field",
                     "ipolatev_grib2", "ipolatev_grib1", "ipolatev_grib2_single_field", "ipolatev_grib1_single_field",
                     "gdswzd", "gdswzd_grib1", "ipgpu_ipolates_grib2_dev", "ipgpu_ipolatev_grib2_dev", "ipxwafs", "ipxwafs2", "ipxwafs3",
                     "ipxetas"):
            assert f"{stem}_{kind}" in names, f"{stem}_{kind} not declared"
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.ipgpu_version.restype = C.c_char_p
    assert b"sm_100a" in lib.ipgpu_version()


def test_header_compiles_as_c():
    src = '#include "ipb200.h"\nint main(void){return 0;}\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_correctly_rounded_functions_match_binary128(tmp_path):
    exe = str(tmp_path / "crmath_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", os.path.join(ROOT, "tests", "native", "crmath_check.cpp"),
                           "-lquadmath", "-o", exe])
    r = subprocess.run([exe, "120000"], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout


@pytest.mark.parametrize("grid,ip", [("203", 6)])
def test_libm_sensitivity_of_the_reference_algorithm(grid, ip):
    """Same algorithm, two libms (glibc 2.39 vs correctly rounded): ipolatev _d results differ by far more than
    1e-12 on a measurable fraction of points, because movect and the rotated grids amplify 1-ulp differences.
    Both flavours still pass the reference's own regression tolerance (test_oracle_golden.py)."""
    kind = "d"
    gin, gout = ic.g2_input(kind, True), ic.g2_templates(kind, True)[grid]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    I = ic.inputs()
    args = (ip, ic.ipopt_for(ip), gin, gout, mi, mo, 1, [0], np.ones(mi, np.uint8), I["u"], I["v"])
    a = oracle(kind, "cr").ipolatev(*args)
    b = oracle(kind, "glibc").ipolatev(*args)
    assert a["iret"] == b["iret"] == 0 and np.array_equal(a["lo"], b["lo"])
    scale = np.abs(a["uo"]).max()
    d = np.abs(a["uo"] - b["uo"]) / scale
    frac = float(np.mean(d > 1e-12))
    print(f"grid {grid} ip {ip}: points moved by libm choice > 1e-12: {frac:.3%}, worst {d.max():.3e} (relative to field scale)")
    assert d.max() < 1e-6   # harmless for any user (the reference's ctest tolerance is 1e-4 absolute) ...
    assert np.mean(a["uo"] == b["uo"]) < 1.0   # ... but not bit-reproducible across libms


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("rowlen,no", [(614, 614 * 428), (669, 669 * 1165), (1799, 1799 * 1059), (33, 33 * 5 + 7), (128, 128), (17, 17 * 9)])
def test_tile_geometry_covers_every_point_once(mode, rowlen, no):
    """Host logic of the point-staged kernels (csrc/ipb_gtiles.cuh): every output point belongs to exactly one
    (tile, row, lane) of each tile shape, and nothing outside the grid is addressed."""
    lib = C.CDLL(pkg.PRODUCT_SO)
    lib.ipgpu_tile_point.restype = C.c_int
    lib.ipgpu_tile_count.restype = C.c_int
    if no > 200000:   # the big grids: check the count, the first and last tiles and a sample in between
        ntile = lib.ipgpu_tile_count(mode, rowlen, no)
        tiles = sorted(set([0, 1, ntile - 2, ntile - 1] + list(range(0, ntile, max(1, ntile // 50)))))
        for t in tiles:
            pts = [lib.ipgpu_tile_point(mode, rowlen, no, t, r, l) for r in range(4) for l in range(32)]
            live = [p for p in pts if p >= 0]
            assert len(live) == len(set(live)) and all(0 <= p < no for p in live)
        assert ntile * 128 >= no
        return
    ntile = lib.ipgpu_tile_count(mode, rowlen, no)
    seen = np.zeros(no, np.int32)
    for t in range(ntile):
        for r in range(4):
            for l in range(32):
                p = lib.ipgpu_tile_point(mode, rowlen, no, t, r, l)
                assert -1 <= p < no
                if p >= 0:
                    seen[p] += 1
    assert seen.min() == 1 and seen.max() == 1


@pytest.mark.parametrize("lsize", ["4", "D", "8"])
def test_fortran_shim_fits_free_form_line_limit(lsize):
    """No Fortran compiler exists in the image, so the shim cannot be compiled here; what can be checked is that no
    statement line exceeds the 132-column free-form limit once cpp has expanded the kind macros (gfortran's default),
    and that every bind(c) name the shim refers to is exported by the library."""
    src = os.path.join(ROOT, "fortran", "ip_mod.F90")
    r = subprocess.run(["cpp", "-P", "-traditional-cpp", f"-DLSIZE={lsize}", src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = [ln for ln in r.stdout.splitlines() if ln.strip() and not ln.lstrip().startswith("!")]
    longest = max(code, key=len)
    assert len(longest) <= 132, (len(longest), longest)
    lib = C.CDLL(pkg.PRODUCT_SO)
    names = set(n for n in re.findall(r"name\s*=\s*'([a-z0-9_]+)'", r.stdout) if re.search(r"_[4d8]$", n))   # imported C symbols
    assert len(names) >= 10, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_fortran_management_module_matches_the_library():
    """fortran/ipb200_gpu.F90 (interfaces of the ipgpu_* management entry points): free-form line limit, every bind(c) name
    exported by the library and declared in include/ipb200.h."""
    src = open(os.path.join(ROOT, "fortran", "ipb200_gpu.F90")).read()
    code = [ln for ln in src.splitlines() if ln.strip() and not ln.lstrip().startswith("!")]
    assert max(len(ln) for ln in code) <= 132
    names = re.findall(r"bind\(c,\s*name='([a-z0-9_]+)'\)", src)
    assert len(names) >= 12 and len(set(names)) == len(names)
    lib = C.CDLL(pkg.PRODUCT_SO)
    declared = set(_declared_symbols())
    for n in names:
        assert hasattr(lib, n), n
        assert n in declared, n


def test_kind_shims_export_the_reference_symbols():
    """libip_4 / libip_d / libip_8 (shims/ip_shim.c): the unsuffixed names the reference's libraries export, for C callers."""
    pkg.build_library()
    here = os.path.dirname(pkg.PRODUCT_SO)
    want = ["gdswzd", "gdswzd_grib1"] + [f"ipolate{sv}_grib{g}{sf}" for sv in "sv" for g in "12" for sf in ("", "_single_field")]
    want += ["ipxwafs_", "ipxwafs2_", "ipxwafs3_", "ipxetas_"]      # external subroutines: the Fortran-mangled names
    for k in "4d8":
        r = subprocess.run(["nm", "-D", "--defined-only", os.path.join(here, f"libip_{k}.so")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        have = set(line.split()[-1] for line in r.stdout.splitlines() if " T " in line)
        assert set(want) <= have, (k, sorted(set(want) - have))
    src = '#include "iplib.h"\nint main(void){return 0;}\n'
    for ls in ("4", "D", "8"):
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f"-DLSIZE={ls}", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                           input=src, text=True, capture_output=True)
        assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("threads", [1, 5])
def test_host_pipeline_helpers_without_a_gpu(threads):
    """ipb_hostcopy.h through ipgpu_selftest_host: the caller's LOGICAL*1 array rebuilt from packed bits, and pitched row
    copies on several host threads (what the host-pointer calls do around the device work)."""
    lib = C.CDLL(pkg.PRODUCT_SO)
    f = lib.ipgpu_selftest_host
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p]
    rng = np.random.default_rng(3)
    for no, mo, kn in ((1, 1, 1), (37, 40, 3), (4099, 4100, 7), (1 << 20, (1 << 20) + 3, 9)):
        truth = (rng.random((kn, no)) < 0.9).astype(np.uint8)
        truth[0] = 1
        nw = (no + 31) // 32
        padded = np.zeros((kn, nw * 32), np.uint8)
        padded[:, :no] = truth
        bits = np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(kn, nw).copy()
        hasfalse = (truth.min(axis=1) == 0).astype(np.int32)
        bits[hasfalse == 0] = 0xDEADBEEF            # rows of fields without a false point are not read
        lo = np.full((kn, mo), 7, np.uint8)
        assert f(0, threads, lo.ctypes.data, mo, bits.ctypes.data, 0, no, kn, hasfalse.ctypes.data) == 0
        assert np.array_equal(lo[:, :no], truth) and (lo[:, no:] == 7).all()
    for width, sp, dp, rows in ((1, 1, 1, 1), (1000, 1024, 1000, 5), (3 << 20, (3 << 20) + 64, (3 << 20) + 8, 4), (5, 9, 7, 700000)):
        src = rng.integers(0, 255, (rows, sp), dtype=np.uint8)
        dst = np.full((rows, dp), 9, np.uint8)
        assert f(1, threads, dst.ctypes.data, dp, src.ctypes.data, sp, width, rows, None) == 0
        assert np.array_equal(dst[:, :width], src[:, :width]) and (dst[:, width:] == 9).all()
    assert f(2, 1, None, 0, None, 0, 0, 0, None) == 1
