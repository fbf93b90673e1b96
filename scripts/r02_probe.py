"""Development probe (not the bench contract): device time of the step for (case x size x knob) combinations, to see what
the kernel's bandwidth depends on.  usage: python scripts/r02_probe.py [quick]   (one B200)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import luma_b200
E = luma_b200


def defs(case, res):
    if case == "cavity":
        return E.Definitions(L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_RE=1000.0, L_UX0=1.0, L_WALL_TOP=E.eVelocity,
                             L_REGULARISED_BOUNDARIES=True, L_NO_FLOW=True)
    if case == "box":        # closed box, all walls solid, no lid: no list kernel at all
        return E.Definitions(L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_RE=1000.0, L_NO_FLOW=True)
    if case.startswith("walls_"):      # walls_xyz: which axes carry solid walls (the others are periodic), e.g. walls_y = the channel;
        ax = case[6:]                  # walls_z4 / walls_z32: z-walls 4 / 32 cells thick (whole sectors / whole warps solid)
        zt = 1
        if ax.startswith("z") and ax[1:].isdigit():
            zt, ax = int(ax[1:]), "z"
        S, F = E.eSolid, E.eFluid
        w = lambda a: S if a in ax else F
        t = lambda a: 1 if a in ax else 0
        return E.Definitions(L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_RE=None, L_NU=1.0 / res, L_NO_FLOW=True,
                             L_WALL_LEFT=w("x"), L_WALL_RIGHT=w("x"), L_WALL_BOTTOM=w("y"), L_WALL_TOP=w("y"), L_WALL_FRONT=w("z"), L_WALL_BACK=w("z"),
                             L_WALL_THICKNESS_CELLS=(t("x"), t("x"), t("y"), t("y"), t("z") * zt, t("z") * zt))
    force = case == "channel_f"
    smag = case == "channel_s"
    return E.Definitions(L_DIMS=3, L_RESOLUTION=res, L_TIMESTEP=0.05 / res, L_RE=None, L_NU=1.0 / res, L_NO_FLOW=True,
                         L_WALL_LEFT=E.eFluid, L_WALL_RIGHT=E.eFluid, L_WALL_FRONT=E.eFluid, L_WALL_BACK=E.eFluid,
                         L_WALL_THICKNESS_CELLS=(0, 0, 1, 1, 0, 0), L_USE_BGKSMAG=smag,
                         L_GRAVITY_ON=force, L_GRAVITY_FORCE=0.0158 if force else 0.0, L_GRAVITY_DIRECTION=0)


def run(case, res, steps, env=None):
    for k in ("LUMA_B200_TMA", "LUMA_B200_V2", "LUMA_B200_STRIDE_PAD", "LUMA_B200_FILL"):
        os.environ.pop(k, None)
    os.environ.update(env or {})
    g = E.GridObj(defs(case, res)).LBM_initGrid()
    g.LBM_multi_opt(10); g.sync()
    best = None
    for _ in range(2):
        g.LBM_multi_opt(steps)
        st = g.stats()
        if best is None or st["ms_per_step"] < best:
            best = st["ms_per_step"]
    g.set_profiling(True)
    g.LBM_multi_opt(20)
    sp = g.stats()
    g.set_profiling(False)
    kms = sp["step_kernel_ms"] / max(1, sp["step_kernel_launches"])
    cells = res ** 3
    print("%-10s %4d^3 %-28s step %.4f ms  %6.0f MLUPS  %5.0f GB/s | k_step alone %.4f ms %5.0f GB/s" % (
        case, res, ",".join("%s=%s" % (k.replace("LUMA_B200_", ""), v) for k, v in (env or {}).items()) or "-",
        best, cells / best / 1e3, cells * 304 / best / 1e6, kms, cells * 304 / kms / 1e6), flush=True)
    g.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "full"
    if mode == "lib":           # one tuning-variant library (LUMA_B200_LIB), the two cases that matter
        print("== library:", os.environ.get("LUMA_B200_LIB", "default"), flush=True)
        for res, st in ((256, 300), (384, 100)):
            for case in ("cavity", "channel_f", "channel_s"):
                run(case, res, st)
        sys.exit(0)
    if mode == "v2":            # the two measured kernel variants against the default kernel
        for res, st in ((256, 300), (384, 100), (512, 40)):
            for case in ("cavity", "box", "channel", "channel_f", "channel_s"):
                run(case, res, st)
                run(case, res, st, {"LUMA_B200_FILL": "1"})
                run(case, res, st, {"LUMA_B200_V2": "1"})
                run(case, res, st, {"LUMA_B200_TMA": "1"})
        sys.exit(0)
    if mode == "walls":         # which wall orientation costs bandwidth?
        for res, st in ((256, 300),):
            for ax in ("", "x", "y", "z", "xy", "xz", "yz", "xyz"):
                run("walls_" + ax, res, st)
                run("walls_" + ax, res, st, {"LUMA_B200_FILL": "1"})
        sys.exit(0)
    if mode == "walls2":        # the sector-completing stores (default) against LUMA_B200_FILL=0
        for res, st in ((256, 300), (384, 100)):
            for case in ("walls_", "walls_z", "walls_z4", "walls_z32", "walls_xyz", "cavity", "channel_f", "channel_s"):
                run(case, res, st)
                run(case, res, st, {"LUMA_B200_FILL": "0"})
        sys.exit(0)
    if mode == "walls3":        # automatic choice per geometry vs forced on / off
        for case in ("cavity", "channel_f", "walls_z", "walls_y"):
            run(case, 256, 300)
            run(case, 256, 300, {"LUMA_B200_FILL": "1"})
            run(case, 256, 300, {"LUMA_B200_FILL": "0"})
        sys.exit(0)
    if mode == "smag":          # the Smagorinsky kernel: periodic channel and the configs[3] geometry at reduced length
        print("== library:", os.environ.get("LUMA_B200_LIB", "default"), flush=True)
        run("channel_s", 256, 300)
        run("channel_s", 384, 100)
        run("cavity", 256, 300)
        d = E.Definitions(L_DIMS=3, L_RESOLUTION=256, L_TIMESTEP=0.05 / 256, L_BX=2.0, L_BY=1.0, L_BZ=1.0, L_RE=7600.0, L_NO_FLOW=True,
                          L_USE_BGKSMAG=True, L_CSMAG=0.3, L_VELOCITY_RAMP=0.5, L_WALL_LEFT=E.eVelocity, L_WALL_RIGHT=E.ePressure,
                          L_WALL_FRONT=E.eFluid, L_WALL_BACK=E.eFluid, L_WALL_THICKNESS_CELLS=(1, 1, 1, 1, 0, 0), body_box=(128, 160, 112, 144, 0, 256))
        g = E.GridObj(d).LBM_initGrid()
        g.LBM_multi_opt(10); g.sync()
        best = None
        for _ in range(2):
            g.LBM_multi_opt(150); st = g.stats()
            best = st["ms_per_step"] if best is None else min(best, st["ms_per_step"])
        cells = 512 * 256 * 256
        print("c4-like 512x256x256 inlet/outlet + cylinder + Smagorinsky: step %.4f ms  %6.0f MLUPS  %5.0f GB/s" % (best, cells / best / 1e3, cells * 304 / best / 1e6), flush=True)
        g.close()
        sys.exit(0)
    if mode == "c1":            # BASELINE configs[0]: 256^2 D2Q9 cavity Re=100, batched call
        d = E.Definitions(L_DIMS=2, L_RESOLUTION=256, L_TIMESTEP=0.05 / 256, L_RE=100.0, L_UX0=1.0, L_WALL_TOP=E.eVelocity,
                          L_REGULARISED_BOUNDARIES=True, L_NO_FLOW=True)
        g = E.GridObj(d).LBM_initGrid()
        g.LBM_multi_opt(200); g.sync()
        for n in (8000, 8000):
            g.LBM_multi_opt(n); st = g.stats()
            print("C1 256^2 D2Q9 cavity, Python batch call of %d steps: %.3f us/step  %.0f MLUPS  graph launches so far %d" % (
                n, 1e3 * st["ms_per_step"], st["mlups_last_call"], st["graph_launches"]), flush=True)
        g.close()
        sys.exit(0)
    if mode == "one":           # a single configuration (for ncu): one <case> <res> [ENV=VALUE ...]
        run(sys.argv[2], int(sys.argv[3]), 20, dict(kv.split("=", 1) for kv in sys.argv[4:]))
        sys.exit(0)
    quick = mode == "quick"
    for res in (256, 384, 512):
        st = {256: 300, 384: 100, 512: 40}[res]
        for case in ("cavity", "box", "channel", "channel_f", "channel_s"):
            run(case, res, st)
            run(case, res, st, {"LUMA_B200_TMA": "1"})
        if quick:
            continue
        for pad in (16, 272, 4112, 65552):
            run("cavity", res, st, {"LUMA_B200_STRIDE_PAD": str(pad)})
        run("channel_f", res, st, {"LUMA_B200_STRIDE_PAD": "4112"})
