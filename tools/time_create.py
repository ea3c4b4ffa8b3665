"""where does engine creation (the e2e leg's first part) spend its time: wraps every C-ABI call with a timer"""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from openems_b200 import load_library

n = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 1024, 1024)
L = load_library()
so, t_build = bench.build_c5(n)
print('pin', so.pin())
op = so.operator()
acc = collections.OrderedDict()


class Wrap:
    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)

    def __getattr__(self, name):
        f = getattr(self._lib, name)
        if not name.startswith("oems_cuda_"):
            return f

        def g(*a):
            t0 = time.perf_counter()
            r = f(*a)
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            return r
        return g


import openems_b200.engine as E
for rep in range(3):
    acc.clear()
    t0 = time.perf_counter()
    eng = op.CreateEngine()
    eng._L = Wrap(eng._L)
    t1 = time.perf_counter()
    bench.add_c5_probes(eng, n)
    eng.Synchronize()
    t2 = time.perf_counter()
    print("rep %d: CreateEngine %.3f s, probes %.3f s" % (rep, t1 - t0, t2 - t1))
    eng.close()
# second pass with the wrapper installed before Init
orig = E.load_library
E.load_library = lambda: Wrap(orig())
for rep in range(2):
    acc.clear()
    t0 = time.perf_counter()
    eng = op.CreateEngine()
    t1 = time.perf_counter()
    print("rep %d: CreateEngine %.3f s; " % (rep, t1 - t0) + ", ".join("%s %.3f" % (k.replace("oems_cuda_", ""), v) for k, v in acc.items() if v > 0.002))
    eng.close()
