"""Quick device-time probe of the step kernel (development aid, not the bench contract)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import luma_b200

def run(res, steps=50, dims=3, smag=False, **kw):
    d = luma_b200.Definitions(L_DIMS=dims, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_RE=1000.0,
                              L_WALL_TOP=luma_b200.eVelocity, L_USE_BGKSMAG=smag, **kw)
    g = luma_b200.GridObj(d).LBM_initGrid()
    g.LBM_multi_opt(5)
    best = 0
    for _ in range(3):
        g.LBM_multi_opt(steps)
        st = g.stats()
        best = max(best, st["mlups_last_call"])
    q = 19 if dims == 3 else 9
    print("res=%d dims=%d smag=%s: %.0f MLUPS  (%.2f ms/step, %.0f GB/s algorithmic)" % (
        res, dims, smag, best, st["ms_per_step"], best * 1e6 * q * 16 / 1e9), flush=True)
    g.close()

if __name__ == "__main__":
    for res in (128, 256, 384):
        run(res)
    run(256, smag=True)
    run(2048, dims=2, steps=200)
