import sys; sys.path.insert(0, '.')
import numpy as np
from tests import cases
from tests.gpu_util import operator_from_oracle
from oracle.pyoracle import BC_PML, BC_PEC
for n, bc, pml in (((23, 26, 21), (BC_PML,) * 6, 4), ((23, 26, 21), (BC_PML, BC_PML, 0, 0, 0, 0), 4), ((23, 26, 21), (0, 0, BC_PML, BC_PML, 0, 0), 4), ((23, 26, 21), (0, 0, 0, 0, BC_PML, BC_PML), 4), ((23, 26, 21), (0, 0,BC_PML, 0, BC_PML, 0), 4)):
    s = cases.uniform_box(n=n, bc=bc, pml=pml)
    eng = operator_from_oracle(s).CreateEngine()
    eng.SetOption("fused", 1)
    if len(sys.argv) > 1:
        eng.SetOption("tma", int(sys.argv[1]))
    print(n, bc, [x for x, _ in eng.TimeSchedule(0)])
    for it in range(30):
        s.iterate(1); eng.IterateTS(1)
        gv, gi = eng.GetFields(0), eng.GetFields(1)
        rv, ri = s.volt, s.curr
        dv = np.argwhere(gv.view(np.uint32) != rv.view(np.uint32))
        di = np.argwhere(gi.view(np.uint32) != ri.view(np.uint32))
        if len(dv) or len(di):
            print(" step", it + 1, "volt diffs", len(dv), dv[:6].tolist(), "curr diffs", len(di), di[:6].tolist())
            break
    else:
        print(" ok 30 steps")
