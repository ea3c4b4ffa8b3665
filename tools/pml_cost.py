"""experiment: cost of the UPML path -- same mesh with PEC, z-only, y-only, x-only and full PML"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for name, bc in (("PEC", [0] * 6), ("PML z", [0, 0, 0, 0, 3, 3]), ("PML y", [0, 0, 3, 3, 0, 0]), ("PML x", [3, 3, 0, 0, 0, 0]), ("PML all", [3] * 6)):
    lines = tuple(np.arange(n, dtype=np.float64) for _ in range(3))
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc(bc, (8,) * 6)
    so.set_excite_gauss(7.5e9, 7.5e9)
    so.add_excitation((n // 2, n // 2, n // 2 + 0.5), (n // 2, n // 2, n // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
    so.build()
    eng = so.CreateEngine()
    for fused in (1, 0):
        eng.SetOption("fused", fused)
        eng.IterateTS(6)
        t = {}
        for k, ms in eng.TimeSchedule(8):
            t[k] = t.get(k, 0) + ms
        st = eng.GetStats()
        ms_graph = eng.IterateTimed(20) / 20
        print("%-8s fused=%d pml_cells %10d  %s  step(events) %.3f ms  step(graph) %.3f ms  %.0f MC/s"
              % (name, fused, st["pml_cells"], " ".join("%s %.3f" % (k, v) for k, v in t.items() if v > 0.02), sum(t.values()),
                 ms_graph, n ** 3 / ms_graph / 1e3), flush=True)
    eng.close()
