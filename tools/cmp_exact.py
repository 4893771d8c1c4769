"""Bit-level comparison of the CUDA library against the oracle on a few _d cases (run on a GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ipcases as ic
from oracle_lib import oracle, product

kind = sys.argv[1] if len(sys.argv) > 1 else "d"
P, O = product(kind), oracle(kind)
R = np.float32 if kind == "4" else np.float64
# 1. projections at fractional coordinates (what the budget sub-samples do)
rng = np.random.default_rng(1)
for name in ("218", "212", "8", "127", "203", "205", "3"):
    g = ic.g2_templates(kind)[name]
    im, jm = g[2][7], g[2][8]
    n = 200000
    x = (rng.random(n) * (im - 1) + 1).astype(R); y = (rng.random(n) * (jm - 1) + 1).astype(R)
    a = P.gdswzd(g, 1, n, xpts=x, ypts=y); b = O.gdswzd(g, 1, n, xpts=x, ypts=y)
    msg = [name, "fwd"]
    for k in ("rlon", "rlat", "crot", "srot", "xlon", "ylat", "area"):
        msg.append(f"{k}:{np.mean(a[k] == b[k]):.5f}")
    a2 = P.gdswzd(g, -1, n, rlon=b["rlon"], rlat=b["rlat"]); b2 = O.gdswzd(g, -1, n, rlon=b["rlon"], rlat=b["rlat"])
    msg.append("inv")
    for k in ("xpts", "ypts", "crot", "srot"):
        msg.append(f"{k}:{np.mean(a2[k] == b2[k]):.5f}")
    print(" ".join(msg), flush=True)
# 2. interpolation
gin = ic.g2_input(kind)
mi = ic.npts_of(gin)
base = ic.inputs()["snoalb"]
km = 2
gi = np.stack([base + k for k in range(km)]).astype(R)
li = (rng.random((km, mi)) < 0.8).astype(np.uint8)
for ip in (0, 1, 3, 6):
    for grid in ("218", "212", "203"):
        gout = ic.g2_templates(kind)[grid]
        mo = ic.npts_of(gout)
        opt = ic.ipopt_for(ip)
        a = P.ipolates(ip, opt, gin, gout, mi, mo, km, [0, 1], li, gi); b = O.ipolates(ip, opt, gin, gout, mi, mo, km, [0, 1], li, gi)
        d = np.abs(a["go"] - b["go"]) / np.maximum(np.abs(b["go"]), 1e-30)
        print("ip", ip, grid, "exact", [float(np.mean(a["go"][k] == b["go"][k])) for k in range(km)], "maxrel", float(d.max()), "lo_eq", bool(np.array_equal(a["lo"], b["lo"])),
              "rlat", float(np.mean(a["rlat"] == b["rlat"])), "rlon", float(np.mean(a["rlon"] == b["rlon"])), flush=True)
