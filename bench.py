#!/usr/bin/env python
"""bench.py — throughput of the ipolates hot path on B200 (BASELINE.json metric: output points x fields / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload bilinear|budget|vector|neighbor|bicubic]
    python bench.py --impl reference ...        # the CPU arm: the oracle port on this box's host cores

One "step" = one ipolates call over the whole field batch (km fields) of the workload:

    bilinear (default; BASELINE configs[1]): ip=0, _4 kind, 0.25-degree global 1440x721 -> 3 km CONUS Lambert
                                              1799x1059, km=1024 synthetic fp32 fields, ibi=0
    budget   (configs[2]): ip=3, _d kind, same grids, km=256, li land-mask bitmap, ibi=1
    vector   (configs[3]): ipolatev ip=0, _4, 0.25-degree global -> NH polar stereographic 512x512, km=512 pairs
    neighbor / bicubic (configs[4]): ip=2 / ip=1, _4, 0.25-degree global -> rotated lat-lon E grid 669x1165, km=1024 per GPU

`value` is measured with every input already resident in HBM (device-pointer C ABI, CUDA events on the
launching stream).  `e2e` is the same metric through the reference-facing host-pointer C ABI
(ipolates_grib2_4 ...) with pinned host buffers: host->device and device->host copies are inside the timed
region.  Multi-GPU: one process per GPU (torchrun), the field batch is sharded — every rank interpolates its
own km fields with the same plan; weak scaling, no data-path collective.

The oracle (oracle/) is executed only for the `cpu_baseline` object and for `--impl reference`.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "ipolates out-pts x fields/s"
UNIT = "pt*fields/s"

# algorithmic bytes per output point x field (derivation in DESIGN.md section 5)
WORKLOADS = {
    "bilinear": dict(ip=0, kind="4", km=1024, src="S", dst="L", vector=False, mask=False,
                     desc="ipolates bilinear, fp32, 0.25deg global 1440x721 -> 3km CONUS Lambert 1799x1059, km=1024"),
    "budget": dict(ip=3, kind="d", km=256, src="S", dst="L", vector=False, mask=True,
                   desc="ipolates budget (ip=3, ipopt=-1) with li bitmap, fp64, 0.25deg global -> 3km CONUS Lambert, km=256"),
    "budget4": dict(ip=3, kind="4", km=256, src="S", dst="L", vector=False, mask=True,
                    desc="ipolates budget (ip=3, ipopt=-1) with li bitmap, fp32 (reference accumulation order, 100 taps per point), 0.25deg global -> 3km CONUS Lambert, km=256"),
    "nbudget": dict(ip=6, kind="4", km=256, src="S", dst="L", vector=False, mask=True,
                    desc="ipolates neighbor-budget (ip=6, ipopt=-1) with li bitmap, fp32, 0.25deg global -> 3km CONUS Lambert, km=256"),
    "vector": dict(ip=0, kind="4", km=512, src="S", dst="P", vector=True, mask=False,
                   desc="ipolatev bilinear u/v, fp32, 0.25deg global -> NH polar stereographic 512x512, km=512 pairs"),
    "neighbor": dict(ip=2, kind="4", km=1024, src="S", dst="E", vector=False, mask=False,
                     desc="ipolates neighbor, fp32, 0.25deg global -> rotated lat-lon E grid 669x1165, km=1024 per GPU"),
    "bicubic": dict(ip=1, kind="4", km=1024, src="S", dst="E", vector=False, mask=False,
                    desc="ipolates bicubic, fp32, 0.25deg global -> rotated lat-lon E grid 669x1165, km=1024 per GPU"),
}


def _pkg():
    return importlib.import_module("nceplibs-ip_b200")


def grids(kind):
    import ipcases as ic
    return ic.baseline_grids(kind)


def ipopt_for(ip):
    return [-1] * 20 if ip in (3, 6) else [0] * 20


def synth_field_np(k, im=1440, jm=721, second=False):
    """gi(i,j,k) of SURVEY.md section 8d (smooth, non-zero, field-dependent)."""
    i = np.arange(im, dtype=np.float64)[None, :]
    j = np.arange(jm, dtype=np.float64)[:, None]
    if second:
        f = 10.0 + 8.0 * np.cos(2 * np.pi * i / im * (1 + k % 5)) * np.sin(np.pi * (j - 360) / jm) - k * 1e-3
    else:
        f = 50.0 + 40.0 * np.sin(2 * np.pi * i / im * (1 + k % 7)) * np.cos(np.pi * (j - 360) / jm) + k * 1e-3
    return f.reshape(-1)


def synth_mask_np(k, im=1440, jm=721):
    """Land-mask-like bitmap: 8x8 blobs, ~70 % true."""
    i = np.arange(im, dtype=np.int64)[None, :]
    j = np.arange(jm, dtype=np.int64)[:, None]
    h = (((i >> 3) * 73856093) ^ ((j >> 3) * 19349663)) % 10
    return (h < 7).astype(np.uint8).reshape(-1)


def survey_plan_bytes_per_point(wl, R_bytes):
    """Plan bytes per output point as SURVEY.md section 8 (a4)/(d) counts them: bilinear NXY(2,2)+WXY(2,2) = 16 + 4 R
    (32 B in the _4 kind), + CXY/SXY/CROT/SROT for vectors, bicubic 16 taps + class byte, neighbor index + x,y; budget: the
    collapsed variant this library applies (9 positions + 9 weights + the weight sum; the survey asks to report the variant)."""
    ip, vec = wl["ip"], wl["vector"]
    if ip == 0:
        return 16 + 4 * R_bytes + (10 * R_bytes if vec else 0)
    if ip == 1:
        return 64 + 16 * R_bytes + 1 + (34 * R_bytes if vec else 0)
    if ip == 2:
        return 4 + 2 * R_bytes
    if ip == 3 and R_bytes == 8:
        return 9 * (4 + R_bytes) + R_bytes
    if ip == 3:
        return 25 * 4 * (4 + R_bytes)   # the _4 kind keeps the reference's order: 25 sub-samples x 4 taps (index + weight)
    return 25 * 4


def algorithmic_bytes(wl, R_bytes, no, touched, plan_bytes_per_pt, km):
    """Compulsory HBM traffic per output point x field: each output element written once, each touched source
    element read once, the plan read once per call."""
    comp = 2 if wl["vector"] else 1
    out_b = comp * R_bytes + 1
    src_b = (comp * R_bytes + (1 if wl["mask"] else 0)) * touched / no
    return out_b + src_b + plan_bytes_per_pt / km


class ClockSampler:
    """SM clock, power and throttle reasons while the GPU is under load (B200_PROFILING.md recipe), sampled every few
    milliseconds through NVML from a side thread — started BEFORE the warm-up steps and stopped after the timed region, so
    that even a 100 ms timed region sees tens of samples; nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None
        self.timed = None       # (t0, t1) of the timed region, perf_counter
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES when it holds plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        try:
            mem = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_MEM))
        except Exception:
            mem = None
        return [time.perf_counter(), float(sm), float(mx), float(pw)] + [bool(r & b) for b in bits] + [mem]

    def _sample_smi(self):
        o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5).stdout.strip()
        x = [v.strip() for v in o.split(",")]
        return [time.perf_counter(), float(x[0]), float(x[1]), float(x[2])] + [v.lower().startswith("active") for v in x[3:7]]

    def _run(self):
        while not self.stop.is_set():
            try:
                self.rows.append(self._sample_nvml() if self.nvml else self._sample_smi())
            except Exception:
                pass
            self.stop.wait(0.002 if self.nvml else 0.05)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        load = self.rows
        if self.timed:
            inside = [r for r in self.rows if self.timed[0] <= r[0] <= self.timed[1]]
        else:
            inside = []
        sm = sorted(r[1] for r in load)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for c, n in enumerate(names) if any(r[4 + c] for r in load)]
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": load[0][2], "power_w_max": max(r[3] for r in load), "samples": len(load),
               "samples_in_timed_region": len(inside), "window": "warm-up + timed region", "source": "nvml" if self.nvml else "nvidia-smi",
               "reasons": reasons}
        if inside:
            smi = sorted(r[1] for r in inside)
            out["sm_mhz_timed_region"] = smi[len(smi) // 2]
        mem = [r[8] for r in (inside or load) if len(r) > 8 and r[8] is not None]
        if mem:   # the HBM clock next to the SM clock: this path is bound by memory, not by the SMs
            out["mem_mhz_min"], out["mem_mhz_max"] = min(mem), max(mem)
            pw = [r[3] for r in (inside or load)]
            out["power_w_first_last"] = [pw[0], pw[-1]]
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port, all host threads, on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------------
def cpu_arm(wl, sample_km=None, target_s=12.0, steps=None, warmup=1):
    """Times the oracle's C entry point itself (buffers preallocated, nothing else inside the timed region)."""
    from oracle_lib import oracle
    kind = wl["kind"]
    # the plain-glibc flavour: the libm a gfortran build of the reference links on this host (oracle/ip_oracle_libm.hpp)
    o = oracle(kind, "glibc")
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what is wanted here)
    try:
        nthr = len(os.sched_getaffinity(0))
    except AttributeError:
        nthr = os.cpu_count() or 1
    o.lib.oracle_set_num_threads(C.c_int(int(os.environ.get("IPB_CPU_THREADS", nthr))))
    o.lib.oracle_num_threads.restype = C.c_int
    cores = int(o.lib.oracle_num_threads())
    R, I = o.R, o.I
    G = grids(kind)
    gin, gout = G[wl["src"]], G[wl["dst"]]
    mi, mo = gin[2][7] * gin[2][8], gout[2][7] * gout[2][8]
    vec = wl["vector"]
    ia = lambda v: np.ascontiguousarray(np.asarray(v, dtype=I).ravel())
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    f = o.fn("ipolatev_grib2" if vec else "ipolates_grib2")

    def prepare(km):
        d = dict(km=km)
        d["gi"] = np.stack([synth_field_np(k) for k in range(km)]).astype(R)
        d["vi"] = np.stack([synth_field_np(k, second=True) for k in range(km)]).astype(R) if vec else None
        d["li"] = np.stack([synth_mask_np(0)] * km) if wl["mask"] else np.ones((1, 1), np.uint8)
        d["go"], d["vo"], d["lo"] = np.zeros((km, mo), R), (np.zeros((km, mo), R) if vec else None), np.zeros((km, mo), np.uint8)
        d["rlat"], d["rlon"], d["crot"], d["srot"] = (np.zeros(mo, R) for _ in range(4))
        d["ibi"], d["ibo"] = ia([1 if wl["mask"] else 0] * km), ia([0] * km)
        sc = dict(ip=ia([wl["ip"]]), opt=ia(ipopt_for(wl["ip"])), numi=ia([gin[1]]), ti=ia(gin[2]), leni=ia([len(gin[2])]), numo=ia([gout[1]]),
                  to=ia(gout[2]), leno=ia([len(gout[2])]), mi=ia([mi]), mo=ia([mo]), km=ia([km]), no=ia([0]), iret=ia([-1]))
        d["sc"] = sc
        head = [P(sc["ip"]), P(sc["opt"]), P(sc["numi"]), P(sc["ti"]), P(sc["leni"]), P(sc["numo"]), P(sc["to"]), P(sc["leno"]),
                P(sc["mi"]), P(sc["mo"]), P(sc["km"]), P(d["ibi"])]
        if vec:
            d["args"] = head + [P(d["li"]), P(d["gi"]), P(d["vi"]), P(sc["no"]), P(d["rlat"]), P(d["rlon"]), P(d["crot"]), P(d["srot"]),
                                P(d["ibo"]), P(d["lo"]), P(d["go"]), P(d["vo"]), P(sc["iret"])]
        else:
            d["args"] = head + [P(d["li"]), P(d["gi"]), P(sc["no"]), P(d["rlat"]), P(d["rlon"]), P(d["ibo"]), P(d["lo"]), P(d["go"]), P(sc["iret"])]
        return d

    def call(d):
        t0 = time.perf_counter()
        f(*d["args"])
        dt = time.perf_counter() - t0
        assert d["sc"]["iret"][0] == 0, int(d["sc"]["iret"][0])
        return dt

    if sample_km is None:
        d = prepare(1)
        call(d)                                  # builds (and caches, where the reference does) the plan
        t1 = call(d)
        sample_km = int(max(1, min(64, 0.25 * target_s / max(t1, 1e-4))))
    d = prepare(sample_km)
    for _ in range(max(warmup, 1)):
        t1 = call(d)
    if steps is None:
        steps = int(max(1, min(50, target_s / max(t1, 1e-4))))
    ts = [call(d) for _ in range(steps)]
    t = sum(ts) / len(ts)
    res = {"go": d["go"], "uo": d["go"], "vo": d["vo"], "lo": d["lo"]}
    return dict(value=mo * sample_km / t, seconds=t, km=sample_km, cores=cores, result=res, mo=mo,
                sample=f"{sample_km} of the workload's fields on the full grids per call, plan cached as the reference's SAVE cache allows "
                       f"(budget family: rebuilt per call, as the reference does); mean of {len(ts)} timed call(s) after warm-up, "
                       f"{cores} OpenMP thread(s), host libm = glibc")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = cpu_arm(wl, target_s=max(2.0, 90.0 / max(args.steps + args.warmup, 1)), steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": c["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": c["seconds"] * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if wl["kind"] == "4" else "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample_km": c["km"], "note": "CPU arm: oracle port of the reference's OpenMP path "
                   "(no Fortran compiler in this image, so the reference itself cannot be built); runs on rank 0 only"},
        "cpu_baseline": {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port", "sample": c["sample"]},
        "e2e": {"value": c["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args, wl):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        # NCCL may print a version banner on stdout while the communicator is created; stdout carries exactly one
        # JSON line, so route fd 1 to stderr until the first collective is through
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local)
    pkg = _pkg()
    kind = wl["kind"]
    lib = pkg.load(kind)
    L = lib.lib
    L.ipgpu_set_device(local)
    if getattr(args, "chunk_mib", 0) > 0:
        L.ipgpu_set_chunk_bytes.argtypes = [C.c_longlong]
        L.ipgpu_set_chunk_bytes(C.c_longlong(args.chunk_mib << 20))
    L.ipgpu_launch_count.restype = C.c_longlong
    L.ipgpu_plan_bytes.restype = C.c_longlong
    L.ipgpu_host_alloc.restype = C.c_void_p
    L.ipgpu_host_alloc.argtypes = [C.c_longlong]
    L.ipgpu_host_free.argtypes = [C.c_void_p]
    R, I = lib.R, lib.I
    tR = torch.float32 if R == np.float32 else torch.float64
    G = grids(kind)
    gin, gout = G[wl["src"]], G[wl["dst"]]
    im, jm = gin[2][7], gin[2][8]
    mi, mo = im * jm, gout[2][7] * gout[2][8]
    km_total = args.km or wl["km"]
    strong = args.scaling == "strong"
    # weak: every rank owns km fields; strong: the km fields of the workload are split over the ranks (rank r owns a contiguous slice)
    km = (km_total * (rank + 1) // world - km_total * rank // world) if strong else km_total
    vec, ip = wl["vector"], wl["ip"]
    opt = ipopt_for(ip)

    # ---- synthetic inputs, generated on the device (same formula as synth_field_np) -----------------------
    k_global0 = (km_total * rank // world) if strong else rank * km   # every rank owns its own slice of the field batch
    ii = torch.arange(im, device=dev, dtype=torch.float64)[None, None, :]
    jj = torch.arange(jm, device=dev, dtype=torch.float64)[None, :, None]
    gi = torch.empty((km, mi), dtype=tR, device=dev)
    vi = torch.empty((km, mi), dtype=tR, device=dev) if vec else None
    for k0 in range(0, km, 32):
        kk = torch.arange(k0, min(k0 + 32, km), device=dev, dtype=torch.float64)[:, None, None]
        kg = kk + k_global0
        f = 50.0 + 40.0 * torch.sin(2 * np.pi * ii / im * (1 + torch.remainder(kg, 7))) * torch.cos(np.pi * (jj - 360) / jm) + kg * 1e-3
        gi[k0:k0 + f.shape[0]] = f.reshape(f.shape[0], -1).to(tR)
        if vec:
            f = 10.0 + 8.0 * torch.cos(2 * np.pi * ii / im * (1 + torch.remainder(kg, 5))) * torch.sin(np.pi * (jj - 360) / jm) - kg * 1e-3
            vi[k0:k0 + f.shape[0]] = f.reshape(f.shape[0], -1).to(tR)
    # the first fields come from the host generator so that the CPU spot check sees bit-identical inputs
    # (device and host sin/cos differ in the last binary64 digit)
    for k in range(min(km, 64)):
        gi[k] = torch.from_numpy(synth_field_np(k + k_global0).astype(R)).to(dev)
        if vec:
            vi[k] = torch.from_numpy(synth_field_np(k + k_global0, second=True).astype(R)).to(dev)
    if wl["mask"]:
        li = torch.from_numpy(synth_mask_np(0)).to(dev)[None, :].repeat(km, 1).contiguous()
    else:
        li = torch.ones((1, 1), dtype=torch.uint8, device=dev)   # never read: every ibi is 0
    go = torch.empty((km, mo), dtype=tR, device=dev)
    vo = torch.empty((km, mo), dtype=tR, device=dev) if vec else None
    lo = torch.empty((km, mo), dtype=torch.uint8, device=dev)
    ibi = np.full(km, 1 if wl["mask"] else 0, dtype=I)
    ibo = np.zeros(km, dtype=I)

    def ia(v):
        return np.ascontiguousarray(np.asarray(v, dtype=I).ravel())

    sc = dict(ip=ia([ip]), opt=ia(opt), numi=ia([gin[1]]), ti=ia(gin[2]), leni=ia([len(gin[2])]), numo=ia([gout[1]]), to=ia(gout[2]),
              leno=ia([len(gout[2])]), mi=ia([mi]), mo=ia([mo]), km=ia([km]), no=ia([0]), iret=ia([-1]))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    D = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    head = [P(sc["ip"]), P(sc["opt"]), P(sc["numi"]), P(sc["ti"]), P(sc["leni"]), P(sc["numo"]), P(sc["to"]), P(sc["leno"]),
            P(sc["mi"]), P(sc["mo"]), P(sc["km"]), P(ibi)]
    if vec:
        fdev = lib.fn("ipgpu_ipolatev_grib2_dev")
        dev_args = [stream] + head + [D(li), D(gi), D(vi), P(sc["no"]), None, None, None, None, P(ibo), D(lo), D(go), D(vo), P(sc["iret"])]
    else:
        fdev = lib.fn("ipgpu_ipolates_grib2_dev")
        dev_args = [stream] + head + [D(li), D(gi), P(sc["no"]), None, None, P(ibo), D(lo), D(go), P(sc["iret"])]

    def step():
        fdev(*dev_args)
        if sc["iret"][0] != 0:
            raise RuntimeError(f"ipolates returned iret={int(sc['iret'][0])}")

    pm, am, hit = C.c_double(0), C.c_double(0), C.c_int(0)
    clk = ClockSampler(local)
    clk.__enter__()                                   # samples from here (plan build, warm-up) to the end of the timed region
    step()                                            # first call builds the plan (reported separately)
    L.ipgpu_last_timing(C.byref(pm), C.byref(am), C.byref(hit))
    plan_build_ms = pm.value
    no = int(sc["no"][0])
    # the same build once more with a warm process (kernels loaded, memory pool grown): what a further grid pair costs
    L.ipgpu_plan_clear()
    step()
    L.ipgpu_last_timing(C.byref(pm), C.byref(am), C.byref(hit))
    plan_rebuild_ms = pm.value
    for _ in range(max(args.warmup - 1, 0)):
        step()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.ipgpu_launch_count()
    apply_ms = []
    barrier()
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
        L.ipgpu_last_timing(C.byref(pm), C.byref(am), C.byref(hit))
        apply_ms.append(am.value)
    e1.record()
    barrier()
    clk.timed = (t_begin, time.perf_counter())
    clk.__exit__()
    ms = e0.elapsed_time(e1)
    launches = L.ipgpu_launch_count() - launches0
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), float(t[1])
    else:
        launches = float(t[1])
    ms_per_step = ms / args.steps
    units_all = no * (km_total if strong else world * km)       # point x fields all ranks processed per step
    value = units_all / (ms_per_step * 1e-3)

    # ---- roofline of the dominant (apply) kernel ------------------------------------------------------
    plan_bytes = L.ipgpu_plan_bytes()
    touched = args_touched(wl)
    Rb = np.dtype(R).itemsize
    # algorithmic bytes per unit exactly as SURVEY.md section 8(d) defines them: output written once, touched source read
    # once, the reference's plan read once per call.  (The plan this library keeps in HBM is larger — staging tables,
    # packed weight records — and mostly not read by the apply kernel; that variant is reported beside it.)
    plan_pp = survey_plan_bytes_per_point(wl, Rb)
    bpu = algorithmic_bytes(wl, Rb, no, touched, plan_pp, max(km, 1))
    bpu_repo = algorithmic_bytes(wl, Rb, no, touched, plan_bytes / max(no, 1), max(km, 1))
    k_ms = sum(apply_ms) / len(apply_ms)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bpu * no * km / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_traffic(args.workload) if not (args.km or args.kind or args.mask >= 0) else None,   # measured at the workload's own size only
                "kernel_ms": k_ms, "kernel_ms_spread": {"min": min(apply_ms), "median": sorted(apply_ms)[len(apply_ms) // 2], "max": max(apply_ms),
                                                        "first_quarter_mean": sum(apply_ms[:max(1, len(apply_ms) // 4)]) / max(1, len(apply_ms) // 4),
                                                        "last_quarter_mean": sum(apply_ms[-max(1, len(apply_ms) // 4):]) / max(1, len(apply_ms) // 4)},
                "bytes_per_unit": bpu, "bytes_per_unit_with_this_librarys_whole_plan": bpu_repo, "units_per_launch": no * km,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, burst)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- the alternative north_star names for N > 1: build once, ncclBroadcast the plan ("measure both") --
    # The plan's arrays stay inside the library; what is timed here is the collective that shipping them would need:
    # one NCCL broadcast of the plan's byte count from rank 0, against the per-rank rebuild this library does.
    plan_bcast_ms = None
    if dist is not None and world > 1:
        buf = torch.empty(int(plan_bytes), dtype=torch.uint8, device=dev)
        for _ in range(2):
            dist.broadcast(buf, src=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(buf, src=0)
        e1.record()
        torch.cuda.synchronize()
        tb = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        plan_bcast_ms = float(tb[0])
        del buf

    # ---- end to end through the host-pointer C ABI (pinned host buffers) --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, lib, L, torch, dist, dev, world, gin, gout, mi, mo, km, no, gi, vi, li, go, vo, lo, ibi, P, ia)

    # ---- CPU baseline (rank 0, single GPU runs only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        c = cpu_arm(wl, target_s=args.cpu_seconds)
        # the sample doubles as a spot check of the device result (same inputs, first fields of the batch)
        kk = min(c["km"], km)
        ref = c["result"]
        got = go[:kk].cpu().numpy()
        lo_g = lo[:kk].cpu().numpy()
        key = "uo" if vec else "go"
        same_lo = bool(np.array_equal(lo_g, ref["lo"][:kk]))
        if kind == "4":
            match = float(np.mean(got.view(np.uint32) == ref[key][:kk].view(np.uint32)))
        else:
            match = float(np.mean(got == ref[key][:kk]))
        scale = float(np.abs(ref[key][:kk]).max()) or 1.0
        cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port", "sample": c["sample"],
               "spot_check": {"fields": kk, "lo_equal": same_lo, "bit_identical_fraction": match,
                              "max_abs_diff_over_field_scale": float(np.abs(got - ref[key][:kk]).max() / scale)}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32" if kind == "4" else "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "km_per_gpu": km, "km_total": km_total if strong else world * km, "no": no, "mi": mi, "ip": ip, "kind": "_" + kind,
                       "cache_defeat": "inputs+outputs per step (%.1f GB) exceed L2 (126 MB)" % ((gi.numel() * (2 if vec else 1) * Rb + go.numel() * (2 if vec else 1) * Rb + lo.numel()) / 1e9),
                       "plan": "cached (built once: %.2f ms in a cold process, %.2f ms warm, %.1f MB in HBM); sharding: km fields per rank, no collective" % (plan_build_ms, plan_rebuild_ms, plan_bytes / 1e6)},
            "plan_build": plan_build_report(args.workload, plan_rebuild_ms, no),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
            "plan_build_ms": plan_build_ms, "plan_rebuild_ms_warm": plan_rebuild_ms,
        }
        if plan_bcast_ms is not None:
            line["plan_nccl_broadcast_ms_same_bytes"] = plan_bcast_ms
        if world == 1 and args.workload == "bilinear" and not args.no_secondary and not args.km:
            line["secondary"] = secondary_budget(args)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def secondary_budget(args):
    """BASELINE's metric names two workloads ("bilinear+budget"); the line above is the bilinear one.  The budget
    configuration (ip=3, _d, land-mask bitmap, km=256) is measured right after it in a child process — the same code path
    as `bench.py --workload budget`, with its own CPU spot check and its own end-to-end number — and attached under
    "secondary" as a complete line of its own."""
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", "budget", "--steps", "10", "--warmup", "3", "--no-secondary",
               "--cpu-seconds", "6"]
        if args.no_cpu:
            cmd.append("--no-cpu")
        if args.no_e2e:
            cmd.append("--no-e2e")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        keep = ("metric", "value", "unit", "ms_per_step", "dtype", "steps", "warmup", "roofline", "cpu_baseline", "e2e", "gpu_launches",
                "plan_build_ms", "plan_rebuild_ms_warm", "plan_build", "clocks")
        out = {"workload": d["config"]["workload"]}
        out.update({k: d.get(k) for k in keep})
        return out
    except Exception as e:  # the headline line must not depend on it
        return {"error": repr(e)[:200]}


# FP64 work of the plan-building kernels, from the committed ncu capture (profiles/plan_fp64.json: thread-level DFMA / DMUL /
# DADD counts of one plan build of the workload).  north_star asks for the plan build "against FP64 peak".
FP64_PEAK_TFLOPS = 37.2   # 148 SMs x 64 FP64 lanes x 2 flop x 1.965 GHz (B200 nominal: 40 TFLOP/s)


def plan_build_report(workload, rebuild_ms, no):
    rep = {"ms_warm": rebuild_ms, "points_per_s": (no / (rebuild_ms * 1e-3)) if rebuild_ms else None, "fp64_peak_tflops": FP64_PEAK_TFLOPS}
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "plan_fp64.json"))).get(workload)
        if d:
            flops = 2 * d["dfma"] + d["dmul"] + d["dadd"]
            rep.update({"fp64_flop": flops, "fp64_kernel_ms": d["kernel_ms"], "fp64_tflops": flops / (d["kernel_ms"] * 1e-3) / 1e12,
                        "fp64_frac_of_peak": flops / (d["kernel_ms"] * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                        "source": "profiles/plan_fp64.json (ncu sass op counts and durations of the plan kernels, one build)"})
    except Exception:
        pass
    return rep


def args_touched(wl):
    # source points referenced by the plan (measured footprints)
    return {"L": 28716, "P": 426680, "E": 298792 if wl["ip"] == 1 else 287525}[wl["dst"]]


def load_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


def run_e2e(args, wl, lib, L, torch, dist, dev, world, gin, gout, mi, mo, km, no, gi, vi, li, go, vo, lo, ibi, P, ia):
    R, I = lib.R, lib.I
    Rb = np.dtype(R).itemsize
    vec = wl["vector"]
    # bound the pinned host footprint per rank (inputs + outputs of one call)
    per_field = (mi + mo) * Rb * (2 if vec else 1) + mo + (mi if wl["mask"] else 0)
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 64 << 30
    budget = min(avail // (2 * max(world, 1)), 24 << 30)
    kme = int(max(1, min(km, budget // per_field)))

    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if getattr(args, "pageable", False):   # what a Fortran caller has: ordinary memory, staged by the library
            return np.zeros(shape, dtype), None
        p = L.ipgpu_host_alloc(C.c_longlong(n))
        if not p:
            raise MemoryError("cudaMallocHost failed")
        buf = (C.c_char * n).from_address(p)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), p

    bufs = []
    h_gi, p_ = pinned((kme, mi), R); bufs.append(p_)
    h_gi[:] = gi[:kme].cpu().numpy()
    h_vi = None
    if vec:
        h_vi, p_ = pinned((kme, mi), R); bufs.append(p_)
        h_vi[:] = vi[:kme].cpu().numpy()
    if wl["mask"]:
        h_li, p_ = pinned((kme, mi), np.uint8); bufs.append(p_)
        h_li[:] = li[:kme].cpu().numpy()
    else:
        h_li = np.ones((1, 1), np.uint8)
    h_go, p_ = pinned((kme, mo), R); bufs.append(p_)
    h_vo = None
    if vec:
        h_vo, p_ = pinned((kme, mo), R); bufs.append(p_)
    h_lo, p_ = pinned((kme, mo), np.uint8); bufs.append(p_)
    h_rlat, h_rlon = np.zeros(mo, R), np.zeros(mo, R)
    h_crot, h_srot = np.zeros(mo, R), np.zeros(mo, R)
    sc = dict(ip=ia([wl["ip"]]), opt=ia(ipopt_for(wl["ip"])), numi=ia([gin[1]]), ti=ia(gin[2]), leni=ia([len(gin[2])]), numo=ia([gout[1]]),
              to=ia(gout[2]), leno=ia([len(gout[2])]), mi=ia([mi]), mo=ia([mo]), km=ia([kme]), no=ia([0]), iret=ia([-1]))
    ibi_e, ibo_e = np.ascontiguousarray(ibi[:kme]), np.zeros(kme, dtype=I)
    head = [P(sc["ip"]), P(sc["opt"]), P(sc["numi"]), P(sc["ti"]), P(sc["leni"]), P(sc["numo"]), P(sc["to"]), P(sc["leno"]),
            P(sc["mi"]), P(sc["mo"]), P(sc["km"]), P(ibi_e)]
    if vec:
        f = lib.fn("ipolatev_grib2")
        a = head + [P(h_li), P(h_gi), P(h_vi), P(sc["no"]), P(h_rlat), P(h_rlon), P(h_crot), P(h_srot), P(ibo_e), P(h_lo), P(h_go), P(h_vo), P(sc["iret"])]
    else:
        f = lib.fn("ipolates_grib2")
        a = head + [P(h_li), P(h_gi), P(sc["no"]), P(h_rlat), P(h_rlon), P(ibo_e), P(h_lo), P(h_go), P(sc["iret"])]

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ndev_lib = 1
    if getattr(args, "devices", 1) > 1 and world == 1:
        # one process, one host-pointer call, the field batch sharded over n GPUs inside the library (ipgpu_set_devices)
        ndev_lib = int(L.ipgpu_set_devices(C.c_int(args.devices)))
    f(*a)   # warm-up (plan is cached already; staging buffers get allocated)
    assert sc["iret"][0] == 0, sc["iret"][0]
    nstep = max(1, min(args.steps, 5))
    sync()
    t0 = time.perf_counter()
    for _ in range(nstep):
        f(*a)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0]) / nstep
    h2d, d2h = C.c_longlong(0), C.c_longlong(0)
    L.ipgpu_last_transfer(C.byref(h2d), C.byref(d2h))
    # result read-back is part of the call: go/lo are host arrays when it returns
    ok = bool(np.array_equal(h_lo[0], lo[0].cpu().numpy())) and bool(np.array_equal(h_go[0].view(np.uint8), go[0].cpu().numpy().view(np.uint8)))
    ok = ok and bool(np.array_equal(h_go[kme - 1].view(np.uint8), go[kme - 1].cpu().numpy().view(np.uint8)))   # last field: another device when sharded
    if ndev_lib > 1:
        L.ipgpu_set_devices(C.c_int(1))
    for p_ in bufs:
        if p_:
            L.ipgpu_host_free(C.c_void_p(p_))
    # the ceiling of this rank's host link, measured in this process with the same pinned allocator (one large copy per
    # direction, then both at once); the call is D2H dominated, so its floor is d2h bytes / d2h rate
    lp = [C.c_double(0), C.c_double(0), C.c_double(0)]
    L.ipgpu_link_probe.argtypes = [C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ipgpu_link_probe(C.c_longlong(256 << 20), C.byref(lp[0]), C.byref(lp[1]), C.byref(lp[2]))
    d2h_b, h2d_b = int(d2h.value) + 2 * mo * Rb + kme * 4, int(h2d.value) + kme * 4
    floor_ms = max(d2h_b / max(lp[1].value, 1e-9), h2d_b / max(lp[0].value, 1e-9)) / 1e6
    # bytes of the whole job (every rank moves the same amount over its own PCIe link)
    return {"value": world * no * kme / dt, "unit": UNIT, "h2d_bytes_per_step": world * h2d_b,
            "d2h_bytes_per_step": world * d2h_b, "ms_per_step": dt * 1e3, "km_per_gpu": kme, "steps": nstep,
            "pcie_gbs_peak": {"h2d": lp[0].value, "d2h": lp[1].value, "both_directions": lp[2].value, "how": "ipgpu_link_probe, 256 MiB pinned copies"},
            "link_floor_ms": floor_ms, "frac_of_link_ceiling": floor_ms / (dt * 1e3), "numa_node_of_gpu": int(L.ipgpu_device_numa_node()),
            "devices_in_library": ndev_lib,
            "host_arrays": "pageable (staged by the library)" if getattr(args, "pageable", False) else "pinned (ipgpu_host_alloc)",
            "api": ("ipolatev_grib2_" if vec else "ipolates_grib2_") + wl["kind"] + " (host pointers, pinned; only the source window the plan references is copied in; lo comes back as one bit per point and is rebuilt by host threads inside the call)",
            "matches_device_path": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bilinear", choices=sorted(WORKLOADS))
    ap.add_argument("--km", type=int, default=0, help="fields per GPU (default: the workload's)")
    ap.add_argument("--kind", default="", choices=["", "4", "d", "8"], help="build kind override (default: the workload's)")
    ap.add_argument("--mask", type=int, default=-1, help="1: every field carries the land-mask bitmap, 0: none (default: the workload's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pageable", action="store_true", help="e2e with ordinary (pageable) host arrays instead of pinned ones")
    ap.add_argument("--no-secondary", action="store_true", help="skip the budget configuration attached to the default (bilinear) line")
    ap.add_argument("--chunk-mib", type=int, default=0, help="staging size per in-flight chunk of the host-pointer path (default: the library's, 768)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--devices", type=int, default=1, help="single process: shard the host-pointer (e2e) call over this many GPUs inside the library")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --km fields per GPU (default; the driver's SCALE run); strong: the workload's km fields split over the GPUs "
                         "(BASELINE config 5: --workload neighbor|bicubic --km 8192 --scaling strong)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 1)
    wl = dict(WORKLOADS[args.workload])
    if args.kind:
        wl["kind"] = args.kind
        wl["desc"] += f" [kind override _{args.kind}]"
    if args.mask >= 0:
        wl["mask"] = bool(args.mask)
        wl["desc"] += f" [bitmap override {args.mask}]"
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_gpu(args, wl)


if __name__ == "__main__":
    main()
