import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openems_b200 import SyntheticOperator
from openems_b200.synthetic import BC_PML, EXC_E_SOFT

def run(n, a, b, zc=0, steps=30, fill=True):
    lines = tuple(np.arange(m, dtype=np.float64) for m in n)
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([BC_PML] * 6, (8,) * 6)
    so.set_excite_gauss(5e9, 5e9)
    so.add_lorentz(tuple(a), tuple(b), eps_fp=(5e9,), eps_tau=(5e-9,), mue_fp=(5e9,), mue_tau=(5e-9,))
    so.add_excitation((10, 10, 10), (n[0] - 11, n[1] - 11, 10), EXC_E_SOFT, (0, 1, 0))
    so.build()
    op = so.operator()
    res = []
    fields = []
    for fused in (1, 0):
        e = op.CreateEngine()
        e.SetOption("fused", fused)
        if zc: e.SetTuning(0, zc, -1)
        if fill: e.FillFields(3)
        e.IterateTS(steps)
        res.append((e.GetOption("fused"), e.FieldDigest()))
        fields.append((e.GetFields(0), e.GetFields(1)))
        e.close()
    dv = np.argwhere(fields[0][0].view(np.uint32) != fields[1][0].view(np.uint32))
    dc = np.argwhere(fields[0][1].view(np.uint32) != fields[1][1].view(np.uint32))
    print(n, a, b, "zc", zc, "steps", steps, "fill", fill, "equal", res[0][1] == res[1][1], "fused", res[0][0], "ndiff", len(dv), len(dc),
          "V first/last", dv[:3].tolist(), dv[-2:].tolist(), "I first", dc[:3].tolist(), flush=True)

for steps in (1, 2, 3):
    run((48, 48, 48), (12, 12, 12), (36, 36, 36), steps=steps)
run((48, 48, 48), (12, 12, 12), (36, 36, 36), zc=16, steps=2)
run((48, 48, 48), (12, 12, 12), (36, 36, 36), zc=48, steps=2)
run((64, 64, 64), (16, 16, 16), (48, 48, 48), steps=2)
run((160, 40, 40), (16, 12, 12), (140, 30, 30), steps=2)
run((48, 48, 48), (12, 12, 12), (36, 36, 36), steps=30, fill=False)
