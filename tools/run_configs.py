"""MCells/s of the BASELINE.json configs C1-C4 on the GPU engine next to the CPU baseline
(oracle sse-compressed multithreaded restatement), same operator, same timestep count.
Run on the GPU box; writes gpurun_out/configs_r01.json.  (C5 is bench.py.)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.pyoracle import OracleSSE
from tests import configs
from tests.gpu_util import operator_from_oracle

out = {}
threads = os.cpu_count() or 1
for name, make, steps in (("C1_parallel_plate_waveguide_21x21x41", configs.c1_parallel_plate_waveguide, 4000),
                          ("C2_msl_notch_filter_120x81x21", configs.c2_msl_notch_filter, 3000),
                          ("C3_patch_antenna_70x70x40_pml8", configs.c3_patch_antenna, 3000),
                          ("C4_drude_block_192^3", lambda: configs.c4_drude_block((192, 192, 192), (48, 144)), 200)):
    r = make()
    s = r[0] if isinstance(r, tuple) else r
    cells = s.N[0] * s.N[1] * s.N[2]
    eng = operator_from_oracle(s).CreateEngine()
    sched = "one-pass" if eng.GetOption("fused") else "two-pass"
    eng.IterateTS(20)
    ms = eng.IterateTimed(steps)
    st = eng.GetStats()
    gpu = cells * steps / (ms * 1e-3) / 1e6
    ms2 = None
    if sched == "one-pass":  # the other schedule, for comparison
        eng.SetOption("fused", 0)
        eng.IterateTS(20)
        ms2 = eng.IterateTimed(steps)
        eng.SetOption("fused", -1)
    best = 0.0
    for th in sorted({1, min(4, threads), threads}):
        cpu = OracleSSE(s, threads=th)
        cpu.iterate(5)
        n = max(20, steps // 10)
        t0 = time.time(); cpu.iterate(n); dt = time.time() - t0
        v = cells * n / dt / 1e6
        if v > best:
            best, best_th = v, th
        cpu.close()
    out[name] = dict(cells=cells, steps=steps, gpu_mcells_s=round(gpu, 1), gpu_us_per_step=round(ms * 1e3 / steps, 2),
                     kernels_per_step=st["kernels_per_step"], schedule=sched,
                     two_pass_us_per_step=None if ms2 is None else round(ms2 * 1e3 / steps, 2), cpu_mcells_s=round(best, 1), cpu_threads=best_th,
                     n_unique=st["n_unique"])
    print(name, out[name], flush=True)
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs_r01.json", "w"), indent=1)
