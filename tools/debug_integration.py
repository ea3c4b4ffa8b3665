import faulthandler, sys, traceback, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
import numpy as np
from oracle import pyref
from oracle.pyref import RefSim, ENGINE_CUDA
from tests import cases
from tests.ref_util import backend, ref_class, assert_operator_equal
def P(*a):
    print(*a, flush=True)
with backend(lambda *args: RefSim(*args, engine=3, threads=3, cuda_lib=(len(sys.argv) > 2))):
    c = cases.engine_cavity()
P("cpu built")
with backend(lambda *args: RefSim(*args, engine=ENGINE_CUDA)):
    g = cases.engine_cavity()
P("gpu built")
try:
    assert_operator_equal(c, g, "x"); P("operator equal")
    c.iterate(3); g.iterate(3); P("iterated")
    v = g.volt; P("volt", np.abs(v).max(), np.array_equal(v, c.volt))
    P("curr eq", np.array_equal(g.curr, c.curr))
    P(c.voltage_integral((5,4,5),(9,4,5)), g.voltage_integral((5,4,5),(9,4,5)))
    P(c.raw_field(0,(16,6,20)), g.raw_field(0,(16,6,20)))
    P(c.energy(), g.energy())
except Exception:
    traceback.print_exc()
order = sys.argv[1] if len(sys.argv) > 1 else "gc"
for ch in order:
    P("closing", ch); (g if ch == "g" else c).close(); P("closed", ch)
