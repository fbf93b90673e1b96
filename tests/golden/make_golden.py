"""TEST INFRASTRUCTURE -- regenerate tests/golden/<case>.json from the compiled reference.

Runs every oracle/_ref/luma_ref_<case> binary (the UNMODIFIED LUMA v1.7.12 sources compiled by
oracle/Makefile; needs /root/reference at build time, so this script only works in the build
container) and records, for the initial state and every snapshot step of the case,

* sha256 digests of the raw little-endian bytes of f, rho, u (reference AoS layout), LatTyp,
  the per-site wall descriptors and the inlet profiles,
* the scalars the reference derived (omega, nu, gravity, rho_out, momentum-exchange force),
* a few probe values (site index -> rho, u, f[0]) so a mismatch can be localised by eye.

The digests are what pins the oracle: tests/test_oracle_pinned.py requires the C restatement
(oracle/luma_oracle.c) to reproduce them bit-for-bit.

usage: python tests/golden/make_golden.py [case ...]
       python tests/golden/make_golden.py --size [case ...]     # oracle.cases.SIZE_CASES: the benchmarked sizes, OpenMP build
       python tests/golden/make_golden.py --size --check-serial c2_128   # also run the serial build and require equal digests
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port  # noqa: E402
from oracle.cases import BENCH_CASES, CASES, SIZE_CASES  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def probes(case, d):
    ns = case.N * case.M * case.K
    ids = sorted(set(int(x) for x in np.linspace(0, ns - 1, 7)))
    out = {}
    for s in ids:
        out[str(s)] = {
            "rho": float(d["rho"][s]).hex(),
            "u": [float(x).hex() for x in d["u"][s * case.dims:(s + 1) * case.dims]],
            "f0": float(d["f"][s * case.Q]).hex(),
        }
    return out


def make(name, size=False, check_serial=False):
    case = BENCH_CASES[name] if size else CASES[name]
    res = port.run_ref_dump(name, case.steps, omp=size)
    if size and check_serial:
        ser = port.run_ref_dump(name, case.steps, omp=False)
        for tag, d in res.items():
            for nm in ("f", "rho", "u"):
                assert digest(d[nm]) == digest(ser[tag][nm]), ("OpenMP build differs from the serial build", name, tag, nm)
        print("golden: %s OpenMP == serial on every snapshot" % name)
        del ser
    g = {"case": name, "doc": case.doc,
         "reference": "cfdemons/LUMA v1.7.12 compiled (oracle/Makefile)" + (", OpenMP build (L_ENABLE_OPENMP)" if size else ""),
         "N": case.N, "M": case.M, "K": case.K, "Q": case.Q, "dims": case.dims, "snapshots": {}}
    init = res["init"]
    g["meta"] = init["meta"]
    g["init"] = {k: digest(init[k]) for k in ("lattyp", "wall", "ux_in", "uy_in", "uz_in", "xpos", "ypos", "zpos")}
    g["type_counts"] = {str(int(t)): int(c) for t, c in zip(*np.unique(init["lattyp"], return_counts=True))}
    for tag, d in res.items():
        g["snapshots"][tag] = {"f": digest(d["f"]), "rho": digest(d["rho"]), "u": digest(d["u"]),
                               "scalars": d["scalars"], "probes": probes(case, d)}
        for nm in ("rho_timeav", "ui_timeav", "uiuj_timeav"):      # L_COMPUTE_TIME_AVERAGED_QUANTITIES cases
            if nm in d:
                g["snapshots"][tag][nm] = digest(d[nm])
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(g, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("golden:", name, list(g["snapshots"]))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--size" in sys.argv:
        for nm in (args or list(SIZE_CASES)):
            make(nm, size=True, check_serial="--check-serial" in sys.argv)
    else:
        for nm in (args or list(CASES)):
            make(nm)
