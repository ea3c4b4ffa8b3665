"""experiment: one-cell-per-thread two-pass kernels (option "small") vs the float4 / z-march kernels over the mesh size"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT
for n in ((21, 21, 41), (70, 70, 40), (120, 81, 21), (96, 96, 96), (128, 128, 128), (160, 160, 160), (200, 200, 200), (256, 256, 256)):
    lines = tuple(np.arange(m, dtype=np.float64) for m in n)
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([3] * 6 if n[0] > 30 else [2] * 6, (8,) * 6)
    so.set_excite_gauss(7.5e9, 7.5e9)
    so.add_excitation((n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), (n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
    so.build()
    eng = so.CreateEngine()
    eng.SetOption("fused", 0)
    steps = 2000 if n[0] * n[1] * n[2] < 3e6 else 300
    res = []
    for small, rows in ((0, 4), (1, 4), (1, 8), (1, 2)):
        eng.SetOption("small", small)
        eng.SetTuning(rows, 0, 1)
        eng.IterateTS(20)
        res.append("small=%d rows=%d: %.2f us" % (eng.GetOption("small"), rows, eng.IterateTimed(steps) / steps * 1e3))
    eng.SetOption("fused", 1)
    eng.IterateTS(20)
    res.append("one-pass: %.2f us" % (eng.IterateTimed(steps) / steps * 1e3))
    print(n, " | ".join(res), flush=True)
    eng.close()
