"""Generates the fixtures of this directory FROM THE ORACLE (oracle/fdtd_oracle.c), not from openEMS:
the reference cannot be built or imported in this container, and its test tree holds no
golden vectors for the time loop.  The fixtures therefore do not pin the oracle against the reference
(tests/test_oracle_pinning.py does what can be done there); they freeze its output so that an accidental
change of the oracle, of a test case builder or of the CUDA engine shows up as a bit difference.

usage: python tests/golden/make_golden.py      (rewrites tests/golden/*.npz)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import cases  # noqa: E402
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
    for name, make, steps, probes in (("cavity_mur_pml_pmc", cavity_case, 240, PROBES),
                                      ("uniform_allpml_40x36x44", allpml_case, 120, [((20, 18, 10), (20, 18, 30)), ((8, 8, 8), (30, 8, 8))])):
        s = make()
        series, digests = record(s, steps, probes)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), series=series, digests=digests, dT=np.float64(s.dT),
                            probes=np.array(probes, np.int64), steps=np.int64(steps))
        print(name, "dT", s.dT, "max|U|", np.abs(series).max(), "digests", digests.tolist())


if __name__ == "__main__":
    main()
