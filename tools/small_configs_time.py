"""quick timing of the tutorial-size configs C1-C3 (us per timestep, kernels per timestep)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import configs
from tests.gpu_util import operator_from_oracle
for name, make in (("C1", configs.c1_parallel_plate_waveguide), ("C2", configs.c2_msl_notch_filter), ("C3", configs.c3_patch_antenna)):
    r = make()
    s = r[0] if isinstance(r, tuple) else r
    eng = operator_from_oracle(s).CreateEngine()
    eng.IterateTS(50)
    us = eng.IterateTimed(4000) / 4000 * 1e3
    cells = s.N[0] * s.N[1] * s.N[2]
    print(name, s.N, "%.2f us/step  %.0f MCells/s" % (us, cells / us), "kernels", [k for k, _ in eng.TimeSchedule(0)], flush=True)
    eng.close()
