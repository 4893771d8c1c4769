"""Small calls through every round-2 kernel path, for compute-sanitizer (memcheck / racecheck / synccheck):
bilinear via the TMA-staged box kernel (whole and edge patches, a bitmap field, a ragged batch), neighbor via the
patch kernel, bicubic and budget via the point-staged kernels, a plan-handle apply.  Exits non-zero on a parity failure."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ipcases as ic  # noqa: E402
from oracle_lib import oracle, product  # noqa: E402

rng = np.random.default_rng(3)
bad = 0
for kind, ip, grid, km in (("4", 0, "218", 11), ("d", 0, "218", 3), ("4", 2, "203", 5), ("4", 1, "8", 9), ("d", 3, "218", 9), ("4", 0, "3", 2)):
    gin, gout = ic.g2_input(kind), ic.g2_templates(kind)[grid]
    mi, mo = ic.npts_of(gin), ic.npts_of(gout)
    R = np.float32 if kind == "4" else np.float64
    g = np.stack([ic.inputs()["snoalb"] + rng.standard_normal(mi) for _ in range(km)]).astype(R)
    li = (rng.random((km, mi)) < 0.8).astype(np.uint8)
    ibi = [1 if k == 1 else 0 for k in range(km)]
    opt = ic.ipopt_for(ip)
    a = product(kind).ipolates(ip, opt, gin, gout, mi, mo, km, ibi, li, g)
    b = oracle(kind).ipolates(ip, opt, gin, gout, mi, mo, km, ibi, li, g)
    same = np.array_equal(a["lo"], b["lo"]) and (np.array_equal(a["go"], b["go"]) if not (ip == 3 and kind == "d") else
                                                 float(np.abs(a["go"] - b["go"]).max()) <= 1e-12 * float(np.abs(b["go"]).max()))
    print(kind, ip, grid, km, "ok" if same else "MISMATCH", flush=True)
    bad += 0 if same else 1
sys.exit(1 if bad else 0)
