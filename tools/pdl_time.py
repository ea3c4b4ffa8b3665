"""experiment: programmatic dependent launch on/off (option "pdl"), headline mesh and a small mesh"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT
for n, steps in (((1024, 1024, 1024), 20), ((70, 70, 40), 3000), ((21, 21, 41), 4000)):
    lines = tuple(np.arange(m, dtype=np.float64) for m in n)
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([3] * 6 if n[0] > 30 else [2] * 6, (8,) * 6)
    so.set_excite_gauss(7.5e9, 7.5e9)
    so.add_excitation((n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), (n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
    so.build()
    eng = so.CreateEngine()
    for rep in range(2):
        for pdl in (1, 0):
            eng.SetOption("pdl", pdl)
            eng.IterateTS(10)
            ms = eng.IterateTimed(steps) / steps
            print(n, "pdl", pdl, "%.3f us/step" % (ms * 1e3), "kernels/step", eng.GetStats()["kernels_per_step"], flush=True)
    eng.close()
