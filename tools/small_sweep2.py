"""experiment: crossover between the one-cell-per-thread two-pass kernels and the one-pass schedule"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT
for n in ((256, 256, 256), (320, 320, 320), (384, 384, 384), (448, 448, 448), (512, 512, 512), (640, 640, 640), (1024, 1024, 512)):
    lines = tuple(np.arange(m, dtype=np.float64) for m in n)
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([3] * 6, (8,) * 6)
    so.set_excite_gauss(7.5e9, 7.5e9)
    so.add_excitation((n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), (n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
    so.build()
    eng = so.CreateEngine()
    steps = 40
    res = []
    for fused, small in ((0, 0), (0, 1), (1, 0)):
        eng.SetOption("small", small)
        eng.SetOption("fused", fused)
        eng.IterateTS(5)
        res.append("fused=%d small=%d: %.1f us" % (eng.GetOption("fused"), eng.GetOption("small"), eng.IterateTimed(steps) / steps * 1e3))
    print(n, " | ".join(res), flush=True)
    eng.close()
