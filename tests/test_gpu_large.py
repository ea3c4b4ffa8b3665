"""GPU parity at the geometries the headline bench runs: many x tiles (nx >= 1024), many y tiles and z chunks
(256^3, PML_8 on all faces) and BASELINE config C4 at its real size (512^3, Drude block 256^3) -- against the
sse-compressed multithreaded restatement of the reference engine (oracle.pyoracle.OracleSSE, itself bit-pinned to
the reference's Engine_SSE_Compressed by tests/test_ref_pinning.py).  Bit-exact E and H."""
import os

import numpy as np
import pytest

from oracle.pyoracle import OracleSim, OracleSSE, BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT
from tests import cases, configs
from tests.gpu_util import operator_from_oracle
from openems_b200 import SyntheticOperator

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 4


def host_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def equal_bits(eng, v, c, what):
    for w, ref in ((0, v), (1, c)):
        got = eng.GetFields(w)
        bad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
        assert bad == 0, "%s: %d values of field %d differ" % (what, bad, w)
        del got


@pytest.mark.parametrize("n", [(1024, 24, 40), (1040, 21, 37), (2051, 19, 20)])
def test_wide_x_many_tiles(n):
    """8+ x tiles of the one-pass kernel (TMA boxes of 136 columns), nx not a multiple of anything"""
    s = cases.uniform_box(n=n, bc=(BC_PML,) * 6, pml=8, src_pos=(n[0] // 2 + 3, n[1] // 2, n[2] // 2))
    op = operator_from_oracle(s)
    sse = OracleSSE(s, THREADS)
    for fused in (1, 0):
        eng = op.CreateEngine()
        eng.SetOption("fused", fused)
        assert eng.GetOption("fused") == fused
        total = 0
        for k in (1, 2, 27):
            eng.IterateTS(k)
            total += k
            if fused:
                sse.iterate(k)
        if not fused:
            pass
        v, c = sse.fields()
        assert sse.num_ts == total
        equal_bits(eng, v, c, "%s fused=%d" % (n, fused))
        assert np.abs(v).max() > 0
        eng.close()


def test_256_cubed_pml8_both_schedules():
    n = (256, 256, 256)
    s = cases.uniform_box(n=n, bc=(BC_PML,) * 6, pml=8)
    op = operator_from_oracle(s)
    sse = OracleSSE(s, THREADS)
    sse.iterate(20)
    v, c = sse.fields()
    assert np.abs(v).max() > 0 and np.abs(c).max() > 0
    for fused in (1, 0):
        eng = op.CreateEngine()
        eng.SetOption("fused", fused)
        assert eng.GetOption("fused") == fused and (not fused or eng.GetOption("tma") == 1)
        eng.IterateTS(20)
        equal_bits(eng, v, c, "256^3 fused=%d" % fused)
        eng.close()


def test_mur_delayed_start_on_gpu():
    """a source ON a Mur plane delays that plane (engine_ext_mur_abc.cpp:44-60): the other planes run from step 0,
    the shared edges follow the plane that is active (run-time resolution of the write order)"""
    lines = tuple(np.arange(m) * 1e-3 for m in (20, 18, 22))
    s = OracleSim(*lines, 1.0)
    s.set_bc([BC_MUR] * 6)
    s.set_excite_gauss(6e9, 6e9)
    s.add_excitation((0.0, 0.004, 0.005), (0.0, 0.012, 0.015), EXC_E_SOFT, (0, 1, 0))
    s.build()
    starts = [m["start_ts"] for m in s.mur_planes()]
    assert starts[0] > 0 and max(starts[1:]) == 0
    from tests.test_gpu_parity import run_both
    run_both(s, steps=(1, 5, starts[0] - 8, 1, 1, 1, 1, 30), what="Mur delayed start")


def test_c4_drude_block_512_cubed():
    """BASELINE config C4 at its real size: 512^3, PML_8 x6, central 256^3 Drude eps+mue block, plane source"""
    if host_gb() < 40:
        pytest.skip("needs ~30 GB of host memory for the dense oracle operator")
    n = (512, 512, 512)
    s = configs.c4_drude_block(n=n, block=(128, 384))
    L = s.lorentz()
    assert len(L) == 1 and L[0]["count"] >= 256 ** 3
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    sse = OracleSSE(s, THREADS)
    steps = 12
    sse.iterate(steps)
    eng.IterateTS(steps)
    v, c = sse.fields()
    assert np.abs(v[:, 128:384, 128:384, 128:384]).max() >= 0 and np.abs(v).max() > 0
    equal_bits(eng, v, c, "C4 512^3")
