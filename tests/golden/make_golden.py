#!/usr/bin/env python
"""Stage the reference's own regression fixtures into tests/golden/ (run once, in the build container).

Source: /root/reference/tests/data/{input_data,baseline_data} (NCEPLIBS-ip 5.0.0 ctest data; fp32
little-endian, column-major, one record per field; vector files hold u then v).  The GPU box has no
/root/reference, so everything the tests need is committed here:

  inputs.npz                        global_snoalb (360x180) and global_uv_wind (360x181, u and v)
  full/<grib>_<kind>_<s|v>_grid<G>.opt<IP>.bin.xz   record 1 (scalar) / records 1-2 (vector), verbatim, xz
  strided.npz                       every STRIDE-th point of the remaining files (key = same stem)

Full copies: grib2 scalar 4- and 8-byte, grib2 vector 8-byte (the exact pins).  The grib1 files and the
grib2 4-byte vector files are kept strided to bound the repo size; STRIDE is stored in the archive.
Spectral (opt4) files are out of scope and skipped.
"""
import lzma, os, sys
import numpy as np

REF = "/root/reference/tests/data"
OUT = os.path.dirname(os.path.abspath(__file__))
STRIDE = 5
NPTS = {"3": 360 * 181, "8": 116 * 44, "127": 768 * 384, "203": 669 * 1165, "205": 954 * 835, "212": 512 * 512, "218": 614 * 428}

def main():
    sno = np.fromfile(f"{REF}/input_data/scalar/global_snoalb.bin", dtype="<f4")
    uv = np.fromfile(f"{REF}/input_data/vector/global_uv_wind.bin", dtype="<f4")
    np.savez_compressed(f"{OUT}/inputs.npz", snoalb=sno, uv=uv)
    os.makedirs(f"{OUT}/full", exist_ok=True)
    strided = {}
    for grib in ("grib1", "grib2"):
        for sv in ("scalar", "vector"):
            for kind, sub in (("4", "4_byte_bin"), ("8", "8_byte_bin")):
                d = f"{REF}/baseline_data/{grib}/{sv}/{sub}"
                for fn in sorted(os.listdir(d)):
                    if ".opt4." in fn:
                        continue
                    grid = fn.split(".")[0][4:]
                    n = NPTS[grid] * (2 if sv == "vector" else 1)
                    a = np.fromfile(f"{d}/{fn}", dtype="<f4")[:n]  # record 1 (+2); two files are over-length
                    assert a.size == n, (fn, a.size, n)
                    stem = f"{grib}_{kind}_{sv[0]}_{fn.rsplit('.', 1)[0]}"
                    full = grib == "grib2" and (sv == "scalar" or kind == "8")
                    if full:
                        with lzma.open(f"{OUT}/full/{stem}.bin.xz", "wb", preset=9) as f:
                            f.write(a.tobytes())
                    else:
                        strided[stem] = a.reshape(-1, NPTS[grid])[:, ::STRIDE].copy()
    np.savez_compressed(f"{OUT}/strided.npz", stride=np.int64(STRIDE), **strided)
    print("wrote", OUT)

if __name__ == "__main__":
    main()
