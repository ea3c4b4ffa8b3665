"""profiling target: a few timesteps of the default (one-pass) schedule, for ncu
usage: prof_fused.py [nx ny nz] [bc0,..,bc5]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT
a = sys.argv[1:]
n = tuple(int(v) for v in a[:3]) if len(a) >= 3 else (1024, 1024, 256)
bc = [int(c) for c in a[-1].split(",")] if a and "," in a[-1] else [3] * 6
lines = tuple(np.arange(m, dtype=np.float64) for m in n)
so = SyntheticOperator(*lines, 1e-3)
so.set_bc(bc, (8,) * 6)
so.set_excite_gauss(7.5e9, 7.5e9)
so.add_excitation((n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), (n[0] // 2, n[1] // 2, n[2] // 2 + 0.5), EXC_E_SOFT, (0, 0, 1))
so.build()
eng = so.CreateEngine()
if os.environ.get("XSLAB") is not None:
    eng.SetOption("xslab", int(os.environ["XSLAB"]))
eng.SetTuning(0, 0, 0)   # no graph: ncu sees plain launches
eng.IterateTS(6)
eng.Synchronize()
print(eng.TimeSchedule(4))
