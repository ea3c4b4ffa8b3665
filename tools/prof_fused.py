import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT
n = (512, 512, 256)
bc = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 0, 0, 0, 3, 3]
lines = tuple(np.arange(m, dtype=np.float64) for m in n)
so = SyntheticOperator(*lines, 1e-3)
so.set_bc(bc, (8,) * 6)
so.set_excite_gauss(7.5e9, 7.5e9)
so.add_excitation((256, 256, 128.5), (256, 256, 128.5), EXC_E_SOFT, (0, 0, 1))
so.build()
eng = so.CreateEngine()
eng.SetTuning(0, 32, 0)
eng.IterateTS(6)
eng.Synchronize()
print(eng.TimeSchedule(4))
