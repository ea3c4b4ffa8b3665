"""experiment: the x-slab variants of the one-pass schedule at the headline mesh
usage: xslab_time.py [nx ny nz]   (default 1024^3, PML_8 on all faces)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT

n = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else [1024, 1024, 1024]
lines = tuple(np.arange(m, dtype=np.float64) for m in n)
so = SyntheticOperator(*lines, 1e-3)
so.set_bc([3] * 6, (8,) * 6)
so.set_excite_gauss(7.5e9, 7.5e9)
so.add_excitation((n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), (n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
so.build()
eng = so.CreateEngine()
for xs, zc in ((0, 0), (2, 16), (2, 32), (2, 64), (2, 128), (1, 0)):
    eng.SetOption("xslab", xs)
    if zc:
        eng.SetOption("xslab_zchunk", zc)
    eng.IterateTS(6)
    t = {}
    for k, ms in eng.TimeSchedule(8):
        t[k] = t.get(k, 0) + ms
    ms_graph = eng.IterateTimed(20) / 20
    print("%s xslab=%d(%d active) zchunk=%d  %s  step(graph) %.3f ms  %.0f MC/s"
          % (n, xs, eng.GetOption("xslab"), zc, " ".join("%s %.3f" % (k, v) for k, v in t.items() if v > 0.02),
             ms_graph, n[0] * n[1] * n[2] / ms_graph / 1e3), flush=True)
eng.close()
