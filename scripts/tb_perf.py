"""Development aid: device time of the two-steps-per-sweep path for several strip/lag/ring settings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import luma_b200

def run(res, geom, steps=101, walls="cavity"):
    kw = dict(L_WALL_TOP=luma_b200.eVelocity) if walls == "cavity" else dict(
        L_WALL_LEFT=1, L_WALL_RIGHT=1, L_WALL_FRONT=1, L_WALL_BACK=1, L_WALL_THICKNESS_CELLS=(0, 0, 1, 1, 0, 0))
    d = luma_b200.Definitions(L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_RE=1000.0, **kw)
    g = luma_b200.GridObj(d).LBM_initGrid()
    if geom is not None:
        g.set_temporal_blocking(True, *geom)
    g.LBM_multi_opt(5)
    best = 0
    for _ in range(3):
        g.LBM_multi_opt(steps)
        st = g.stats()
        best = max(best, st["mlups_last_call"])
    print("res=%d %s geom=%s: %.0f MLUPS (%.3f ms/step) [%s]" % (res, walls, geom, best, st["ms_per_step"],
          g.temporal_blocking_status() if geom is not None else "one-step"), flush=True)
    g.close()

if __name__ == "__main__":
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    run(res, None)
    for geom in [(0, 0, 0), (32, 4, 7), (64, 4, 7), (128, 4, 7), (256, 3, 6), (64, 2, 5), (64, 6, 9), (32, 8, 11), (16, 8, 11), (128, 2, 5), (96, 3, 6)]:
        try:
            run(res, geom)
        except Exception as e:
            print("geom", geom, "failed:", e, flush=True)
