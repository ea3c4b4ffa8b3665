"""Tuning sweep of the TMA-staged one-pass kernel (development aid): tile rows, ring depth, z chunk.
usage: python tools/sweep_tma.py build   (CPU box: cross-compiles the variants)
       python tools/sweep_tma.py run     (GPU box)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "ty7_s3": [],
    "ty7_s2": ["FT_STAGES=2"],
    "ty15_s3": ["FUSED_TY=15", "FT_MIN_BLOCKS=1"],
    "ty15_s2": ["FUSED_TY=15", "FT_STAGES=2", "FT_MIN_BLOCKS=1"],
    "ty5_s3": ["FUSED_TY=5"],
    "ty3_s3": ["FUSED_TY=3", "FT_MIN_BLOCKS=3"],
    "ty3_s4": ["FUSED_TY=3", "FT_STAGES=4", "FT_MIN_BLOCKS=3"],
}
LIBDIR = os.path.join(ROOT, "openems_b200", "lib", "variants")


def build():
    from openems_b200 import build as b
    os.makedirs(LIBDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        out = os.path.join(LIBDIR, "lib_%s.so" % name)
        b.build(force=True, defines=defs, out=out)
        print("built", out)


def worker():
    import bench
    n = 1024
    so, _ = bench.build_c5((n, n, n))
    eng = so.operator().CreateEngine()
    res = []
    for zc in (6, 8, 12, 16, 24):
        eng.SetTuning(0, zc, 0)
        eng.IterateTS(3)
        eng.Synchronize()
        t = dict()
        for name, ms in eng.TimeSchedule(5):
            t[name] = t.get(name, 0) + ms
        res.append((zc, eng.GetOption("tma"), round(t["fused_EH"], 4), round(sum(t.values()), 4)))
    print("RESULT " + json.dumps(res))


def run():
    for name in VARIANTS:
        lib = os.path.join(LIBDIR, "lib_%s.so" % name)
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, OPENEMS_B200_LIB=lib)
        out = subprocess.run([sys.executable, __file__, "worker"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        line = [l for l in out.splitlines() if l.startswith("RESULT ")]
        print(name, line[0][7:] if line else out[-800:], flush=True)


if __name__ == "__main__":
    {"build": build, "worker": worker, "run": run}[sys.argv[1]]()
