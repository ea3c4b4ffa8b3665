"""experiment: cost of the UPML shell of the one-pass schedule for one mesh / boundary set
usage: shell_cost.py nx ny nz bc0,bc1,bc2,bc3,bc4,bc5"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT

n = [int(a) for a in sys.argv[1:4]]
bc = [int(c) for c in sys.argv[4].split(",")]
lines = tuple(np.arange(m, dtype=np.float64) for m in n)
so = SyntheticOperator(*lines, 1e-3)
so.set_bc(bc, (8,) * 6)
so.set_excite_gauss(7.5e9, 7.5e9)
so.add_excitation((n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), (n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
so.build()
eng = so.CreateEngine()
if os.environ.get("XSLAB") is not None:
    eng.SetOption("xslab", int(os.environ["XSLAB"]))
for fused, tma in ((1, 1), (1, 0), (0, 0)):
    eng.SetOption("fused", fused)
    eng.SetOption("tma", tma)
    eng.IterateTS(6)
    t = {}
    for k, ms in eng.TimeSchedule(8):
        t[k] = t.get(k, 0) + ms
    st = eng.GetStats()
    ms_graph = eng.IterateTimed(20) / 20
    print("%s bc=%s fused=%d tma=%d pml_cells %d  %s  step(graph) %.3f ms  %.0f MC/s"
          % (n, sys.argv[4], fused, eng.GetOption("tma"), st["pml_cells"], " ".join("%s %.3f" % (k, v) for k, v in t.items() if v > 0.02),
             ms_graph, n[0] * n[1] * n[2] / ms_graph / 1e3), flush=True)
eng.close()
