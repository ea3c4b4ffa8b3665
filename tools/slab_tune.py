"""experiment: launch shape for a 1024x1024x128 slab-sized mesh (what each GPU sees at N=8)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 128
so, _ = bench.build_c5((1024, 1024, nz))
eng = so.operator().CreateEngine()
for rows, zc in ((4, 32), (4, 16), (4, 8), (4, 4), (2, 16), (2, 8), (8, 8), (8, 16)):
    eng.SetTuning(rows, zc, 0)
    eng.IterateTS(3)
    t = {}
    for k, ms in eng.TimeSchedule(10):
        t[k] = t.get(k, 0) + ms
    eng.SetTuning(rows, zc, 1)
    ms_graph = eng.IterateTimed(50) / 50
    print("rows %d zchunk %2d  E %.4f  H %.4f  sum %.4f   graph step %.4f ms  %.0f MC/s" % (rows, zc, t["update_E"], t["update_H"], sum(t.values()), ms_graph, 1024 * 1024 * nz / ms_graph / 1e3), flush=True)
