"""Generates the fixtures of this directory FROM THE REFERENCE ITSELF: oracle/_ref/libopenems_ref.so, i.e. the
unmodified translation units of /root/reference (operator, multithreaded sse-compressed engine -- the
reference's default engine --, extensions, Engine_Interface_FDTD) compiled by oracle/Makefile.ref and driven
by oracle/ref_driver.cpp.  Needs /root/reference (or the prebuilt library).  Each fixture holds the timestep,
voltage-probe series of every timestep and bit digests of E and H at three timesteps.

The CPU suite checks the oracle restatement against them (tests/test_oracle_pinning.py), the GPU suite the
CUDA engine, both schedules (tests/test_gpu_parity.py).

usage: python tests/golden/make_golden.py      (rewrites tests/golden/*.npz)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import cases, configs  # noqa: E402
from oracle.pyoracle import BC_PML, BC_MUR, BC_PEC, BC_PMC  # noqa: E402

PROBES = [((5, 4, 5), (6, 4, 5)), ((10, 2, 20), (10, 8, 20)), ((20, 5, 12), (20, 5, 28))]


def field_digest(a):
    """order-sensitive 64-bit digest of a float32 array's bit patterns"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64).ravel()
    w = (np.arange(u.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)) | np.uint64(1)
    return np.uint64(np.bitwise_xor.reduce((u + np.uint64(0x632BE5AB)) * w))


def cavity_case():
    return cases.engine_cavity()


def allpml_case():
    return cases.uniform_box(n=(40, 36, 44), bc=(BC_PML,) * 6, pml=8)


def ppw_sinus_case():
    return configs.c1_parallel_plate_waveguide("sinus")[0]


def drude_case():
    return configs.c4_drude_block(n=(30, 30, 30), block=(10, 20))


def patch_case():
    return configs.c3_patch_antenna(n=(40, 40, 30))[0]


CASES = {
    "cavity_mur_pml_pmc": (cavity_case, 240, PROBES),
    "uniform_allpml_40x36x44": (allpml_case, 120, [((20, 18, 10), (20, 18, 30)), ((8, 8, 8), (30, 8, 8))]),
    "ppw_sinus_mur_pmc": (ppw_sinus_case, 150, [((10, 0, 25), (10, 20, 25)), ((10, 0, 35), (10, 20, 35))]),
    "drude_block_30": (drude_case, 100, [((5, 15, 15), (25, 15, 15)), ((15, 12, 8), (15, 18, 8))]),
    "patch_lumped_rc_pml": (patch_case, 80, [((17, 20, 12), (17, 20, 14)), ((10, 10, 16), (30, 10, 16))]),
}


def record(s, steps, probes):
    series = np.zeros((steps, len(probes)), np.float64)
    digests = []
    for t in range(steps):
        s.iterate(1)
        for q, (a, b) in enumerate(probes):
            series[t, q] = s.voltage_integral(a, b)
        if (t + 1) in (1, 10, steps):
            digests.append((t + 1, field_digest(s.volt), field_digest(s.curr)))
    return series, np.array(digests, np.uint64)


def main():
    from oracle import pyref
    from tests.ref_util import backend, ref_class
    for name, (make, steps, probes) in CASES.items():
        with backend(ref_class(pyref.ENGINE_MULTITHREADED, 3)):
            s = make()
        series, digests = record(s, steps, probes)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), series=series, digests=digests, dT=np.float64(s.dT),
                            probes=np.array(probes, np.int64), steps=np.int64(steps),
                            source=np.bytes_(b"oracle/_ref/libopenems_ref.so (" + pyref.lib().ref_version() + b"), Engine_Multithread"))
        print(name, "dT", s.dT, "max|U|", np.abs(series).max(), "digests", digests.tolist())


if __name__ == "__main__":
    main()
