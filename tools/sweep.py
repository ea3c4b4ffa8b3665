"""Kernel tuning sweep (development aid, run on the GPU box): builds variants of the library
with different tuning macros and times the stencil kernels for several launch shapes.
usage: python tools/sweep.py build   (on the CPU box: cross-compiles the variants)
       python tools/sweep.py run N   (on the GPU box)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "mb4": [],
    "pf1": ["OEMS_PREFETCH_DIST=1"],
    "pf2": ["OEMS_PREFETCH_DIST=2"],
    "pf4": ["OEMS_PREFETCH_DIST=4"],
    "pf8": ["OEMS_PREFETCH_DIST=8"],
    "pf2_all": ["OEMS_PREFETCH_DIST=2", "OEMS_PREFETCH_LANES=0"],
    "pf2_mb3": ["OEMS_PREFETCH_DIST=2", "OEMS_MIN_BLOCKS=3"],
}
LIBDIR = os.path.join(ROOT, "openems_b200", "lib", "variants")


def build():
    from openems_b200 import build as b
    os.makedirs(LIBDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        out = os.path.join(LIBDIR, "lib_%s.so" % name)
        b.build(force=True, defines=defs, out=out)
        print("built", out)


def worker(n, shapes):
    import numpy as np
    import bench
    so, _ = bench.build_c5((n, n, n))
    eng = so.operator().CreateEngine()
    res = []
    for rows, zc in shapes:
        eng.SetTuning(rows, zc, 0)
        eng.IterateTS(3)
        eng.Synchronize()
        t = dict()
        for name, ms in eng.TimeSchedule(5):
            t[name] = t.get(name, 0) + ms
        res.append((rows, zc, round(t["update_E"], 4), round(t["update_H"], 4)))
    print("RESULT " + json.dumps(res))


def run(n):
    shapes = [(8, 32), (4, 32), (2, 32), (2, 64), (1, 32)]
    for name in VARIANTS:
        lib = os.path.join(LIBDIR, "lib_%s.so" % name)
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, OPENEMS_B200_LIB=lib)
        out = subprocess.run([sys.executable, __file__, "worker", str(n), json.dumps(shapes)], env=env,
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        line = [l for l in out.splitlines() if l.startswith("RESULT ")]
        print(name, line[0][7:] if line else out[-500:])


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "worker":
        worker(int(sys.argv[2]), json.loads(sys.argv[3]))
    else:
        run(int(sys.argv[2]))
