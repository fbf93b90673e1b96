import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import luma_b200
from oracle import port
from oracle.cases import CASES
from util import defs_from_case
name = sys.argv[1]; geom = tuple(int(x) for x in sys.argv[2].split(","))
case = CASES[name]
ref = port.PortGrid(case)
g = luma_b200.GridObj(defs_from_case(case)).LBM_initGrid()
g.set_temporal_blocking(True, *geom)
print(g.temporal_blocking_status())
done = 0
for s in (1, 2, 3, 4, 11):
    g.LBM_multi_opt(s - done); ref.step(s - done); done = s
    got = g.download()
    for nm, wdt in (("f", case.Q), ("rho", 1), ("u", case.dims)):
        a = got[nm].reshape(-1, wdt); b = getattr(ref, nm).reshape(-1, wdt)
        bad = np.flatnonzero((a != b).any(axis=1))
        if bad.size:
            ids = bad[:8]
            MK = case.M * case.K
            print("t=%d %s: %d sites differ; first (i,j,k): %s types %s" % (s, nm, bad.size,
                  [(int(i // MK), int((i % MK) // case.K), int(i % case.K)) for i in ids], ref.lattyp[ids]))
            js = sorted(set(int((i % MK) // case.K) for i in bad)); print("   rows j:", js[:40]); 
            ps = sorted(set(int(i // MK) for i in bad)); print("   planes:", ps[:40])
            break
    else:
        print("t=%d ok" % s)
        continue
    break
