"""experiment: one-pass vs two-pass schedule over mesh sizes (where is the crossover?)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT

for n in [int(a) for a in sys.argv[1:]] or [48, 64, 96, 128, 192, 256, 384, 512]:
    lines = tuple(np.arange(n, dtype=np.float64) for _ in range(3))
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([3] * 6, (8,) * 6)
    so.set_excite_gauss(7.5e9, 7.5e9)
    so.add_excitation((n // 2, n // 2, n // 2 + 0.5), (n // 2, n // 2, n // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
    so.build()
    eng = so.CreateEngine()
    steps = max(20, min(2000, int(2e9 / n ** 3)))
    res = []
    for fused, zc in ((0, 0), (1, 0), (1, 8), (1, 4), (1, 2)):
        eng.SetOption("fused", fused)
        eng.SetTuning(0, zc, -1)
        eng.IterateTS(10)
        ms = eng.IterateTimed(steps) / steps
        res.append("%s zc=%d: %.1f us (%.0f MC/s)" % ("one-pass" if fused else "two-pass", zc, ms * 1e3, n ** 3 / ms / 1e3))
    print(n, " | ".join(res), flush=True)
    eng.close()
