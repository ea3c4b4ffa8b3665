"""experiment: where the time of the small configs (C1-C3) goes, and block rows / z-chunk of the two-pass kernels"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import configs
from tests.gpu_util import operator_from_oracle
for name, make in (("C1", configs.c1_parallel_plate_waveguide), ("C2", configs.c2_msl_notch_filter), ("C3", configs.c3_patch_antenna)):
    r = make()
    s = r[0] if isinstance(r, tuple) else r
    eng = operator_from_oracle(s).CreateEngine()
    eng.IterateTS(50)
    print(name, s.N, "schedule (us):", [(k, round(ms * 1e3, 2)) for k, ms in eng.TimeSchedule(50)], flush=True)
    for rows in (4, 8, 2):
        for zc in (1, 2, 4, 8):
            eng.SetTuning(rows, zc, 1)
            eng.IterateTS(50)
            us = eng.IterateTimed(2000) / 2000 * 1e3
            print("   rows %d zchunk %d: %.2f us/step" % (rows, zc, us), flush=True)
    eng.close()
