"""GPU tests of the device-pointer entry points (ipgpu_ipolates_grib2_dev_<k>) — the path `bench.py` times.

The headline kernel shifts its thread -> point mapping per field so that every vector store is aligned, which
depends on the alignment of the caller's go / lo pointers and on k*mo mod 4.  These tests drive it with output
arrays that start at every element offset 0..3, with field counts that are not multiples of 4 or of the staging
pass, with and without bitmaps, and compare bit for bit with the oracle.
"""
import ctypes as C

import numpy as np
import pytest

import ipcases as ic
from oracle_lib import oracle, product

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev_call(lib, ip, opt, gin, gout, mi, mo, km, ibi, li_t, gi_t, go_t, lo_t):
    I = lib.I
    ia = lambda v: np.ascontiguousarray(np.asarray(v, dtype=I).ravel())
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    sc = dict(ip=ia([ip]), opt=ia(opt), numi=ia([gin[1]]), ti=ia(gin[2]), leni=ia([len(gin[2])]), numo=ia([gout[1]]), to=ia(gout[2]),
              leno=ia([len(gout[2])]), mi=ia([mi]), mo=ia([mo]), km=ia([km]), no=ia([0]), iret=ia([-1]))
    ibi_a, ibo_a = ia(ibi), ia([0] * km)
    D = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.fn("ipgpu_ipolates_grib2_dev")(stream, P(sc["ip"]), P(sc["opt"]), P(sc["numi"]), P(sc["ti"]), P(sc["leni"]), P(sc["numo"]), P(sc["to"]),
                                       P(sc["leno"]), P(sc["mi"]), P(sc["mo"]), P(sc["km"]), P(ibi_a), D(li_t), D(gi_t), P(sc["no"]), None, None,
                                       P(ibo_a), D(lo_t), D(go_t), P(sc["iret"]))
    torch.cuda.synchronize()
    return int(sc["iret"][0]), int(sc["no"][0]), ibo_a


@pytest.mark.parametrize("kind", ["4", "d"])
@pytest.mark.parametrize("km,off", [(1, 0), (3, 1), (5, 2), (9, 3), (13, 1), (37, 0)])
@pytest.mark.parametrize("masked", [False, True])
def test_bilinear_device_api_alignment_and_ragged_batches(kind, km, off, masked):
    lib = product(kind)
    R = lib.R
    tR = torch.float32 if R == np.float32 else torch.float64
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)["218"]      # 1 degree -> 12 km Lambert: the tile-staged kernel
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    rng = np.random.default_rng(100 + km + off)
    base = ic.inputs()["snoalb"]
    gi = np.stack([base * (1 + 0.01 * k) + rng.standard_normal(mi) * (k % 3) for k in range(km)]).astype(R)
    gi[0, ::7] = 0.0                                                 # exact zeros go through the window test of the fast division
    if km > 1:
        gi[1] *= 1e-30                                               # sums below the exact-residual window: IEEE division fallback
    li = (rng.random((km, mi)) < 0.85).astype(np.uint8)
    ibi = [(k % 2) if masked else 0 for k in range(km)]
    dev = torch.device("cuda", 0)
    gi_t = torch.from_numpy(gi).to(dev)
    li_t = torch.from_numpy(li).to(dev)
    go_all = torch.full((km * mo + 8,), -7.0, dtype=tR, device=dev)
    lo_all = torch.full((km * mo + 8,), 9, dtype=torch.uint8, device=dev)
    go_t, lo_t = go_all[off:off + km * mo], lo_all[off:off + km * mo]   # same element phase for go and lo
    iret, no, ibo = _dev_call(lib, 0, [0] * 20, gin, gout, mi, mo, km, ibi, li_t, gi_t, go_t, lo_t)
    ref = oracle(kind).ipolates(0, [0] * 20, gin, gout, mi, mo, km, ibi, li, gi)
    assert iret == ref["iret"] == 0 and no == ref["no"] == mo
    got = go_t.cpu().numpy().reshape(km, mo)
    lo = lo_t.cpu().numpy().reshape(km, mo)
    assert np.array_equal(lo, ref["lo"])
    assert np.array_equal(ibo, ref["ibo"])
    assert np.array_equal(got.view(np.uint8), ref["go"].view(np.uint8)), float(np.abs(got - ref["go"]).max())
    # nothing outside the output range was touched
    assert float(go_all[:off].sum()) == -7.0 * off and bool((go_all[off + km * mo:] == -7.0).all())
    assert bool((lo_all[:off] == 9).all()) and bool((lo_all[off + km * mo:] == 9).all())


def test_mismatched_go_lo_phase_uses_general_kernel():
    """go and lo starting at different element phases cannot share aligned vector stores: results must still be exact."""
    kind = "4"
    lib = product(kind)
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)["218"]
    mi, mo, km = ic.npts_of(gin), ic.npts_of(gout), 3
    gi = np.stack([ic.inputs()["snoalb"] + k for k in range(km)]).astype(np.float32)
    dev = torch.device("cuda", 0)
    gi_t = torch.from_numpy(gi).to(dev)
    li_t = torch.ones((1,), dtype=torch.uint8, device=dev)
    go_all = torch.zeros((km * mo + 8,), dtype=torch.float32, device=dev)
    lo_all = torch.zeros((km * mo + 8,), dtype=torch.uint8, device=dev)
    go_t, lo_t = go_all[1:1 + km * mo], lo_all[2:2 + km * mo]
    iret, no, ibo = _dev_call(lib, 0, [0] * 20, gin, gout, mi, mo, km, [0] * km, li_t, gi_t, go_t, lo_t)
    ref = oracle(kind).ipolates(0, [0] * 20, gin, gout, mi, mo, km, [0] * km, np.ones((km, mi), np.uint8), gi)
    assert iret == 0 and no == mo
    assert np.array_equal(go_t.cpu().numpy().reshape(km, mo).view(np.uint32), ref["go"].view(np.uint32))
    assert np.array_equal(lo_t.cpu().numpy().reshape(km, mo), ref["lo"])


@pytest.mark.parametrize("km", [1, 7, 40])
def test_budget_binary64_collapsed_path_lo_exact_values_1e12(km):
    """_d budget with land-mask bitmaps through the collapsed kernel: lo / ibo bit-exact, values within 1e-12."""
    kind = "d"
    lib = product(kind)
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)["218"]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    rng = np.random.default_rng(5 + km)
    gi = np.stack([ic.inputs()["snoalb"] * (1 + 0.1 * k) + rng.standard_normal(mi) for k in range(km)]).astype(np.float64)
    jj, ii = np.divmod(np.arange(mi), 360)
    li = np.stack([((((ii >> 2) * 73856093) ^ ((jj >> 2) * 19349663) ^ (k * 83492791)) % 10 < 6).astype(np.uint8) for k in range(km)])
    ibi = [1 if k % 3 else 0 for k in range(km)]
    opt = [-1] * 20
    a = lib.ipolates(3, opt, gin, gout, mi, mo, km, ibi, li, gi)
    b = oracle(kind).ipolates(3, opt, gin, gout, mi, mo, km, ibi, li, gi)
    assert a["iret"] == b["iret"] == 0 and a["no"] == b["no"]
    assert np.array_equal(a["lo"], b["lo"]) and np.array_equal(a["ibo"], b["ibo"])
    scale = np.abs(b["go"]).max()
    assert float(np.abs(a["go"] - b["go"]).max()) <= 1e-12 * scale
