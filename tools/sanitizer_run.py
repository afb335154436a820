import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import __graft_entry__ as e
pkg = e.load_package()
for which, n in ((2, 6000), (3, 6000), (4, 3000), (1, 4000)):
    b = pkg.SceneBuilder(which, 128 if which == 4 else 0)
    sc = pkg.Scene(b)
    tu = pkg.TraceUnit(0, 128, 96, seed=5, batch=n)
    pl = pkg.PlotUnit(0, 128, 96)
    tu.render_fused(sc, pl, 0, n)
    ph = tu.render_range(sc, 0, n)
    pl.plot(tu)
    g = pkg.GatherUnit(128, 96); g.accumulate(pl, clear=True)
    t = pkg.TonemapUnit(128, 96); t.tonemap(g)
    print(which, "ok", int(np.count_nonzero(ph["probability"])), tu.ray_count())
