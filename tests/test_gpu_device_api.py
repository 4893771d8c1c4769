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


# ---- explicit plan handles, the non-draining apply, GRIB1 device variants, per-device state (round 2) ----------------
def _ia(lib, v):
    return np.ascontiguousarray(np.asarray(v, dtype=lib.I).ravel())


def _P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("kind", ["4", "d", "8"])
@pytest.mark.parametrize("ip,vector", [(0, False), (2, False), (3, False), (1, False), (0, True), (6, True)])
def test_plan_handle_async_apply(kind, ip, vector):
    """ipgpu_plan_get / geometry / apply (queued twice on one stream, no synchronisation in between) / release."""
    lib = product(kind)
    R = lib.R
    tR = torch.float32 if R == np.float32 else torch.float64
    gin, gout = ic.g2_input(kind, vector), ic.g2_templates(kind, vector)["218"]
    mi, mo, km = ic.npts_of(gin), ic.npts_of(gout), 5
    rng = np.random.default_rng(200 + ip)
    base = ic.inputs()["u" if vector else "snoalb"]
    gi = np.stack([base * (1 + 0.01 * k) + rng.standard_normal(mi) for k in range(km)]).astype(R)
    vi = np.stack([rng.standard_normal(mi) * 3 - k for k in range(km)]).astype(R)
    li = (rng.random((km, mi)) < 0.85).astype(np.uint8)
    ibi = [k % 2 for k in range(km)]
    opt = ic.ipopt_for(ip)
    dev = torch.device("cuda", 0)
    sc = dict(vec=_ia(lib, [int(vector)]), ip=_ia(lib, [ip]), opt=_ia(lib, opt), numi=_ia(lib, [gin[1]]), ti=_ia(lib, gin[2]), leni=_ia(lib, [len(gin[2])]),
              numo=_ia(lib, [gout[1]]), to=_ia(lib, gout[2]), leno=_ia(lib, [len(gout[2])]), mi=_ia(lib, [mi]), mo=_ia(lib, [mo]), no=_ia(lib, [0]),
              iret=_ia(lib, [-1]), km=_ia(lib, [km]))
    get = lib.fn("ipgpu_plan_get")
    get.restype = C.c_void_p
    h = get(_P(sc["vec"]), _P(sc["ip"]), _P(sc["opt"]), _P(sc["numi"]), _P(sc["ti"]), _P(sc["leni"]), _P(sc["numo"]), _P(sc["to"]), _P(sc["leno"]),
            _P(sc["mi"]), _P(sc["mo"]), _P(sc["no"]), _P(sc["iret"]))
    assert h and int(sc["iret"][0]) == 0 and int(sc["no"][0]) == mo
    gi_t, vi_t, li_t = torch.from_numpy(gi).to(dev), torch.from_numpy(vi).to(dev), torch.from_numpy(li).to(dev)
    outs = []
    stream = torch.cuda.Stream()
    sp = C.c_void_p(stream.cuda_stream)
    D = lambda t: C.c_void_p(t.data_ptr())
    apply_ = lib.fn("ipgpu_plan_apply")
    torch.cuda.synchronize()
    for rep in range(2):   # two applies in flight on the same stream, different output arrays
        go_t = torch.full((km, mo), -7.0, dtype=tR, device=dev)
        vo_t = torch.full((km, mo), -7.0, dtype=tR, device=dev)
        lo_t = torch.full((km, mo), 9, dtype=torch.uint8, device=dev)
        ibo_t = torch.full((km,), -3, dtype=torch.int32, device=dev)
        ir = _ia(lib, [-1])
        apply_(C.c_void_p(h), sp, _P(sc["km"]), _P(_ia(lib, ibi)), D(li_t), D(gi_t), D(vi_t) if vector else None, D(ibo_t), D(lo_t), D(go_t),
               D(vo_t) if vector else None, _P(ir))
        assert int(ir[0]) == 0
        outs.append((go_t, vo_t, lo_t, ibo_t))
    geo = {k: torch.zeros(mo, dtype=tR, device=dev) for k in ("rlat", "rlon", "crot", "srot")}
    ir = _ia(lib, [-1])
    lib.fn("ipgpu_plan_geometry")(C.c_void_p(h), sp, D(geo["rlat"]), D(geo["rlon"]), D(geo["crot"]) if vector else None,
                                  D(geo["srot"]) if vector else None, _P(ir))
    stream.synchronize()
    lib.lib.ipgpu_plan_release(C.c_void_p(h))
    O = oracle(kind)
    ref = O.ipolatev(ip, opt, gin, gout, mi, mo, km, ibi, li, gi, vi) if vector else O.ipolates(ip, opt, gin, gout, mi, mo, km, ibi, li, gi)
    for go_t, vo_t, lo_t, ibo_t in outs:
        assert np.array_equal(lo_t.cpu().numpy(), ref["lo"])
        assert np.array_equal(ibo_t.cpu().numpy().astype(np.int64), np.asarray(ref["ibo"]).astype(np.int64))
        names = (("uo", go_t), ("vo", vo_t)) if vector else (("go", go_t),)
        for nm, t in names:
            got = t.cpu().numpy()
            if ip == 3 and kind != "4" and not vector:
                assert float(np.abs(got - ref[nm]).max()) <= 1e-12 * float(np.abs(ref[nm]).max())
            else:
                assert np.array_equal(got.view(np.uint8), ref[nm].view(np.uint8)), (nm, float(np.abs(got - ref[nm]).max()))
    assert np.array_equal(geo["rlat"].cpu().numpy(), ref["rlat"]) and np.array_equal(geo["rlon"].cpu().numpy(), ref["rlon"])
    if vector:
        assert np.array_equal(geo["crot"].cpu().numpy(), ref["crot"]) and np.array_equal(geo["srot"].cpu().numpy(), ref["srot"])


@pytest.mark.parametrize("kind", ["4", "d"])
def test_grib1_device_variants(kind):
    lib = product(kind)
    R = lib.R
    tR = torch.float32 if R == np.float32 else torch.float64
    dev = torch.device("cuda", 0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    D = lambda t: C.c_void_p(t.data_ptr())
    km = 3
    for vector in (False, True):
        gin, gout = ic.g1_input(vector), ic.g1_templates()["218"]
        mi, mo = ic.npts_of(gin), ic.npts_of(gout)
        rng = np.random.default_rng(300)
        gi = np.stack([ic.inputs()["u" if vector else "snoalb"] + k for k in range(km)]).astype(R)
        vi = np.stack([rng.standard_normal(mi) + k for k in range(km)]).astype(R)
        li = np.ones((km, mi), np.uint8)
        gi_t, vi_t, li_t = torch.from_numpy(gi).to(dev), torch.from_numpy(vi).to(dev), torch.from_numpy(li).to(dev)
        go_t = torch.zeros((km, mo), dtype=tR, device=dev)
        vo_t = torch.zeros((km, mo), dtype=tR, device=dev)
        lo_t = torch.zeros((km, mo), dtype=torch.uint8, device=dev)
        sc = dict(ip=_ia(lib, [0]), opt=_ia(lib, [0] * 20), ki=_ia(lib, gin[1]), ko=_ia(lib, gout[1]), mi=_ia(lib, [mi]), mo=_ia(lib, [mo]),
                  km=_ia(lib, [km]), ibi=_ia(lib, [0] * km), no=_ia(lib, [0]), ibo=_ia(lib, [0] * km), iret=_ia(lib, [-1]))
        if vector:
            lib.fn("ipgpu_ipolatev_grib1_dev")(stream, _P(sc["ip"]), _P(sc["opt"]), _P(sc["ki"]), _P(sc["ko"]), _P(sc["mi"]), _P(sc["mo"]), _P(sc["km"]),
                                               _P(sc["ibi"]), D(li_t), D(gi_t), D(vi_t), _P(sc["no"]), None, None, None, None, _P(sc["ibo"]), D(lo_t),
                                               D(go_t), D(vo_t), _P(sc["iret"]))
            ref = oracle(kind).ipolatev(0, [0] * 20, gin, gout, mi, mo, km, [0] * km, li, gi, vi)
            assert np.array_equal(vo_t.cpu().numpy().view(np.uint8), ref["vo"].view(np.uint8))
            key = "uo"
        else:
            lib.fn("ipgpu_ipolates_grib1_dev")(stream, _P(sc["ip"]), _P(sc["opt"]), _P(sc["ki"]), _P(sc["ko"]), _P(sc["mi"]), _P(sc["mo"]), _P(sc["km"]),
                                               _P(sc["ibi"]), D(li_t), D(gi_t), _P(sc["no"]), None, None, _P(sc["ibo"]), D(lo_t), D(go_t), _P(sc["iret"]))
            ref = oracle(kind).ipolates(0, [0] * 20, gin, gout, mi, mo, km, [0] * km, li, gi)
            key = "go"
        torch.cuda.synchronize()
        assert int(sc["iret"][0]) == 0 and int(sc["no"][0]) == mo
        assert np.array_equal(go_t.cpu().numpy().view(np.uint8), ref[key].view(np.uint8))
        assert np.array_equal(lo_t.cpu().numpy(), ref["lo"])


def test_calls_from_other_host_threads_use_their_current_device():
    """The runtime state is per device and picked by the device that is current when a call enters: a call from a
    freshly started host thread (OpenMP worker, Python thread) lands on a consistent set of streams and plans."""
    import threading
    lib = product("4")
    gin, gout = ic.g2_input("4"), ic.g2_templates("4")["218"]
    mi, mo, km = ic.npts_of(gin), ic.npts_of(gout), 4
    g = np.stack([ic.inputs()["snoalb"] + k for k in range(km)]).astype(np.float32)
    li = np.ones((km, mi), np.uint8)
    main = lib.ipolates(0, [0] * 20, gin, gout, mi, mo, km, [0] * km, li, g)
    res = {}

    def work(i):
        res[i] = lib.ipolates(0, [0] * 20, gin, gout, mi, mo, km, [0] * km, li, g)

    th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in range(3):
        assert res[i]["iret"] == 0 and np.array_equal(res[i]["go"], main["go"]) and np.array_equal(res[i]["lo"], main["lo"])


def test_pinned_allocator_and_link_probe():
    lib = product("4").lib
    lib.ipgpu_host_alloc.restype = C.c_void_p
    lib.ipgpu_host_alloc.argtypes = [C.c_longlong]
    lib.ipgpu_host_free.argtypes = [C.c_void_p]
    p = lib.ipgpu_host_alloc(C.c_longlong(8 << 20))
    assert p
    buf = (C.c_ubyte * (8 << 20)).from_address(p)
    buf[0] = 7
    buf[(8 << 20) - 1] = 9
    lib.ipgpu_host_free(C.c_void_p(p))
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    lib.ipgpu_link_probe.argtypes = [C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ipgpu_link_probe(C.c_longlong(64 << 20), C.byref(a), C.byref(b), C.byref(c))
    assert a.value > 1.0 and b.value > 1.0 and c.value > 1.0, (a.value, b.value, c.value)
    print("link GB/s h2d, d2h, both:", a.value, b.value, c.value, "numa node", lib.ipgpu_device_numa_node())


def test_host_register_pins_a_callers_array():
    """ipgpu_host_register: an ordinary array becomes directly copyable; results do not depend on how the arrays are held."""
    P = product("4")
    gin, gout = ic.g2_input("4"), ic.g2_templates("4")["218"]
    mi, mo, km = ic.npts_of(gin), ic.npts_of(gout), 3
    g = np.stack([ic.inputs()["snoalb"] + k for k in range(km)]).astype(np.float32)
    a = P.ipolates(0, [-1] * 20, gin, gout, mi, mo, km, [0] * km, np.ones((km, mi), np.uint8), g)
    assert P.lib.ipgpu_host_register(C.c_void_p(g.ctypes.data), C.c_longlong(g.nbytes)) == 0
    try:
        b = P.ipolates(0, [-1] * 20, gin, gout, mi, mo, km, [0] * km, np.ones((km, mi), np.uint8), g)
    finally:
        assert P.lib.ipgpu_host_unregister(C.c_void_p(g.ctypes.data)) == 0
    assert P.lib.ipgpu_host_register(None, C.c_longlong(16)) != 0
    for k in ("go", "lo", "ibo"):
        assert np.array_equal(a[k], b[k]), k


def test_host_call_sharded_over_devices():
    """ipgpu_set_devices(n): one host-pointer call, the field batch split over n GPUs (needs >= 2 visible devices)."""
    lib = product("4")
    n = lib.lib.ipgpu_device_count()
    if n < 2:
        pytest.skip("one visible device")
    gin, gout = ic.g2_input("4"), ic.g2_templates("4")["218"]
    mi, mo, km = ic.npts_of(gin), ic.npts_of(gout), 11
    rng = np.random.default_rng(9)
    g = np.stack([ic.inputs()["snoalb"] + rng.standard_normal(mi) for k in range(km)]).astype(np.float32)
    li = (rng.random((km, mi)) < 0.8).astype(np.uint8)
    ibi = [k % 2 for k in range(km)]
    one = lib.ipolates(0, [0] * 20, gin, gout, mi, mo, km, ibi, li, g)
    assert lib.lib.ipgpu_set_devices(n) == n
    try:
        many = lib.ipolates(0, [0] * 20, gin, gout, mi, mo, km, ibi, li, g)
        u = np.stack([ic.inputs()["u"][:mi] + k for k in range(km)]).astype(np.float32)
        v = np.stack([ic.inputs()["v"][:mi] - k for k in range(km)]).astype(np.float32)
        vm = lib.ipolatev(0, [0] * 20, ic.g2_input("4", True), ic.g2_templates("4", True)["212"], ic.npts_of(ic.g2_input("4", True)), 512 * 512, km,
                          ibi, np.ones((km, ic.npts_of(ic.g2_input("4", True))), np.uint8), np.resize(u, (km, ic.npts_of(ic.g2_input("4", True)))),
                          np.resize(v, (km, ic.npts_of(ic.g2_input("4", True)))))
    finally:
        lib.lib.ipgpu_set_devices(1)
    for k in ("go", "lo", "ibo", "rlat", "rlon"):
        assert np.array_equal(one[k], many[k]), k
    assert many["iret"] == 0 and many["no"] == mo and vm["iret"] == 0
    v1 = lib.ipolatev(0, [0] * 20, ic.g2_input("4", True), ic.g2_templates("4", True)["212"], ic.npts_of(ic.g2_input("4", True)), 512 * 512, km,
                      ibi, np.ones((km, ic.npts_of(ic.g2_input("4", True))), np.uint8), np.resize(u, (km, ic.npts_of(ic.g2_input("4", True)))),
                      np.resize(v, (km, ic.npts_of(ic.g2_input("4", True)))))
    for k in ("uo", "vo", "lo", "ibo", "crot", "srot"):
        assert np.array_equal(v1[k], vm[k]), k


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("kind", ["4", "d"])
@pytest.mark.parametrize("ip", [0, 2, 3])
def test_host_call_lo_crosses_the_link_packed(kind, ip, pinned):
    """Host-pointer calls bring lo back as one bit per point and rebuild the caller's LOGICAL*1 array on the host
    (k_pack_lo / expand_lo): fields without a false point, fields with holes, mo > no (the tail stays untouched), a point
    count that is no multiple of 8 or 32, and many chunks in flight one after the other.  pinned=False: ordinary (pageable)
    arrays, which the library stages through its own page-locked buffers with host threads; True: direct copies."""
    P, O = product(kind), oracle(kind)
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)["218"]
    mi, no, km = ic.npts_of(gin), ic.npts_of(gout), 9
    mo = no + 37
    rng = np.random.default_rng(31 + ip)
    g = np.stack([ic.inputs()["snoalb"] + rng.standard_normal(mi) for k in range(km)]).astype(P.R)
    li = (rng.random((km, mi)) < 0.7).astype(np.uint8)
    li[4] = 1                                    # a field that announces a bitmap and has no hole
    ibi = [0, 1, 0, 1, 1, 0, 1, 1, 0]
    P.lib.ipgpu_set_chunk_bytes(C.c_longlong(1))   # one field per chunk
    try:
        a = P.ipolates(ip, [-1] * 20, gin, gout, mi, mo, km, ibi, li, g, pinned=pinned)
    finally:
        P.lib.ipgpu_set_chunk_bytes(C.c_longlong(768 << 20))
    b = O.ipolates(ip, [-1] * 20, gin, gout, mi, mo, km, ibi, li, g)
    assert a["iret"] == 0 and a["no"] == b["no"] == no
    assert np.array_equal(a["lo"], b["lo"]) and np.array_equal(a["ibo"], b["ibo"])
    assert (a["lo"][:, no:] == 7).all()          # beyond no the caller's bytes are left alone
    assert (a["go"][:, no:] == -7777.0).all()
    if ip != 3 or kind == "4":
        assert np.array_equal(a["go"].view(np.uint8), b["go"].view(np.uint8))
    else:
        assert np.allclose(a["go"], b["go"], rtol=1e-12, atol=1e-12)
    assert a["lo"][0].min() == 1 and a["lo"][1].min() == 0


@pytest.mark.parametrize("lsize,sfx", [("4", "4"), ("D", "d"), ("8", "8")])
def test_c_caller_links_the_kind_shims(lsize, sfx, tmp_path):
    """A C program written against the reference's C interface (unsuffixed gdswzd / gdswzd_grib1 / ipolates_grib2, include/iplib.h)
    links libip_<kind>.so and passes the checks of tests/tst_gdswzd_grib2.c / tst_gdswzd_grib1.c on the GPU."""
    import importlib
    import os
    import subprocess
    pkg = importlib.import_module("nceplibs-ip_b200")
    pkg.build_library()
    here = os.path.dirname(pkg.PRODUCT_SO)
    root = os.path.dirname(here)
    exe = str(tmp_path / f"tst_gdswzd_{sfx}")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", f"-DLSIZE={lsize}", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "native", "tst_gdswzd_c.c"),
                           "-o", exe, "-L", here, f"-lip_{sfx}", "-lipb200", "-lm", f"-Wl,-rpath,{here}"])
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
