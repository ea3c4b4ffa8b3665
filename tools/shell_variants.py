"""experiment: occupancy / march length of the shell kernels at 1024^3 (variants built with SHELL_MIN_BLOCKS=n)"""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import numpy as np
    from openems_b200 import SyntheticOperator
    from openems_b200.synthetic import EXC_E_SOFT
    n = (1024, 1024, 1024)
    lines = tuple(np.arange(m, dtype=np.float64) for m in n)
    so = SyntheticOperator(*lines, 1e-3)
    so.set_bc([3] * 6, (8,) * 6)
    so.set_excite_gauss(7.5e9, 7.5e9)
    so.add_excitation((512, 512, 512.5), (512, 512, 512.5), EXC_E_SOFT, (0, 0, 1))
    so.build()
    eng = so.CreateEngine()
    for zc in (4, 8, 16, 32, 64):
        eng.SetOption("shell_zchunk", zc)
        eng.IterateTS(4)
        t = {}
        for k, ms in eng.TimeSchedule(6):
            t[k] = t.get(k, 0) + ms
        print("  shell_zchunk %2d: shell_E %.3f shell_H %.3f  step %.3f" % (zc, t["shell_E"], t["shell_H"], sum(t.values())), flush=True)
else:
    for v in ("", "sh2", "sh4", "sh5", "sh6"):
        env = dict(os.environ)
        if v:
            env["OPENEMS_B200_LIB"] = os.path.join(ROOT, "openems_b200", "lib", "variants", "lib_%s.so" % v)
        print("variant", v or "default (3 blocks)", flush=True)
        subprocess.run([sys.executable, __file__, "worker"], env=env)
