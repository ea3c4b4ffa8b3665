"""engine creation at the headline size with OEMS_TIMING=1 (stages of the upload / finalize), three times"""
import sys, os, time
os.environ["OEMS_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = (1024, 1024, 1024)
so, t_build = bench.build_c5(n)
op = so.operator()
for rep in range(3):
    t0 = time.perf_counter()
    eng = op.CreateEngine()
    eng.Synchronize()
    t1 = time.perf_counter()
    print("create %d: %.3f s" % (rep, t1 - t0), flush=True)
    eng.IterateTS(2)
    eng.Synchronize()
    t2 = time.perf_counter()
    eng.close()
    print("  close: %.3f s" % (time.perf_counter() - t2), flush=True)
