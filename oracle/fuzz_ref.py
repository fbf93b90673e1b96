"""TEST INFRASTRUCTURE -- pins the C restatement against the COMPILED reference on seeded random cases (the generator of
tests/test_kernels_fuzz_cpu.py), beyond the fixed case table: every case is written as a definitions header, the
unmodified LUMA sources are compiled with it (oracle/Makefile `one`, ~10 s each), run, and compared bit for bit with
oracle/luma_oracle.c.  Needs /root/reference, so it only runs in the build container:

    python oracle/fuzz_ref.py [first_seed [count]]

Headers, objects and binaries of these cases are scratch (oracle/cases/fuzz*.h is git-ignored)."""
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import port  # noqa: E402
from test_kernels_fuzz_cpu import random_case  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ok = skipped = 0
    for seed in range(first, first + count):
        case = random_case(seed)
        hdr = os.path.join(HERE, "cases", case.name + ".h")
        with open(hdr, "w") as fh:
            fh.write(case.header())
        try:
            subprocess.run(["make", "-C", HERE, "-j8", "one", "CASE=" + case.name, "OMP=0"], check=True, stdout=subprocess.DEVNULL)
            try:
                g = port.PortGrid(case)
            except RuntimeError:
                g = None
            try:
                res = port.run_ref_dump(case.name, case.steps)
            except subprocess.CalledProcessError:
                res = None
            if g is None or res is None:
                # both must refuse (omega >= 2, or an L_ERROR while stepping)
                if res is None and g is not None:
                    try:
                        g.step(max(case.steps))
                        raise SystemExit("seed %d: the reference stops but the port runs on" % seed)
                    except RuntimeError:
                        pass
                elif res is not None:
                    raise SystemExit("seed %d: the port refuses a case the reference runs" % seed)
                skipped += 1
                print("seed %d: refused by both" % seed, flush=True)
                continue
            assert np.array_equal(res["init"]["lattyp"], g.lattyp), seed
            assert np.array_equal(res["init"]["f"], g.f), seed
            for s in case.steps:
                g.step(s - g.t)
                d = res["t%d" % s]
                for nm in ("f", "rho", "u") + (("rho_timeav", "ui_timeav", "uiuj_timeav") if case.time_averaged else ()):
                    a, b = d[nm], getattr(g, nm)
                    same = (a == b) | (np.isnan(a) & np.isnan(b))
                    if not same.all():
                        raise SystemExit("seed %d t=%d %s: %d values differ (%s)" % (seed, s, nm, int((~same).sum()), case))
            ok += 1
            print("seed %d: %dD Q%d %dx%dx%d identical at steps %s" % (seed, case.dims, case.Q, case.N, case.M, case.K, list(case.steps)), flush=True)
            g.close()
        finally:
            os.unlink(hdr)
            shutil.rmtree(os.path.join(HERE, "_build", case.name), ignore_errors=True)
            exe = os.path.join(HERE, "_ref", "luma_ref_" + case.name)
            if os.path.exists(exe):
                os.unlink(exe)
    print("fuzz_ref: %d identical, %d refused by both, seeds %d..%d" % (ok, skipped, first, first + count - 1))


if __name__ == "__main__":
    main()
