import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import EXC_E_SOFT
n = (256, 256, 256)
lines = tuple(np.arange(m, dtype=np.float64) for m in n)
so = SyntheticOperator(*lines, 1e-3)
so.set_bc([3] * 6, (8,) * 6)
so.set_excite_gauss(7.5e9, 7.5e9)
so.add_excitation((128, 128, 128.5), (128, 128, 128.5), EXC_E_SOFT, (0, 0, 1))
so.build()
eng = so.CreateEngine()
print("fused", eng.GetOption("fused"), "tma", eng.GetOption("tma"), "skip", eng.GetOption("skip_shell"))
for sk in (1, 0):
    eng.SetOption("skip_shell", sk)
    eng.IterateTS(4)
    print(sk, eng.GetOption("skip_shell"), eng.TimeSchedule(6))
