"""helpers for the oracle-vs-reference pinning tests: run one case builder on the CPU restatement
(oracle.pyoracle.OracleSim) and on the reference's own classes (oracle.pyref.RefSim) and compare
every operator table, the fields and the read-outs bit for bit."""
import contextlib
import functools

import numpy as np

from oracle import pyref
from oracle.pyoracle import OracleSim
from tests import cases, configs


@contextlib.contextmanager
def backend(cls):
    """case builders in tests/cases.py and tests/configs.py construct `OracleSim`; swap the class"""
    saved = cases.OracleSim, configs.OracleSim
    cases.OracleSim = configs.OracleSim = cls
    try:
        yield
    finally:
        cases.OracleSim, configs.OracleSim = saved


def ref_class(engine=pyref.ENGINE_BASIC, threads=1):
    return functools.partial(pyref.RefSim, engine=engine, threads=threads)


def build_both(fn, *args, engine=pyref.ENGINE_BASIC, threads=1, **kw):
    """-> (oracle result, reference result) of the same case builder"""
    o = fn(*args, **kw)
    with backend(ref_class(engine, threads)):
        r = fn(*args, **kw)
    return o, r


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == np.float32:
        return bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    if a.dtype == np.float64:
        return bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))
    return bool(np.array_equal(a, b))


def assert_same(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    if not bits_equal(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        bad = np.argwhere(d > 0)
        raise AssertionError("%s: %d of %d values differ, max abs diff %g (ref max %g), first at %s"
                             % (what, len(bad), a.size, d.max(), np.abs(b).max(), bad[:4].tolist()))


def assert_operator_equal(o, r, what=""):
    """every table Operator_CUDA::CreateEngine reads: dT, vv/vi/ii/iv, signal, excitation lists, UPML, Mur,
    Lorentz, TFSF, absorbing sheets"""
    assert o.dT == r.dT, "%s dT %r vs %r" % (what, o.dT, r.dT)
    assert o.nyquist == r.nyquist
    for w in ("vv", "vi", "ii", "iv"):
        assert_same(o.coeff(w), r.coeff(w), what + " " + w)
    so, sr = o.signal(), r.signal()
    assert_same(so[0], sr[0], what + " sig_v")
    assert_same(so[1], sr[1], what + " sig_i")
    assert so[2] == sr[2], "%s signal period %d vs %d" % (what, so[2], sr[2])
    for c in (0, 1):
        eo, er = o.excitation(c), r.excitation(c)
        for a, b, n in zip(eo, er, ("idx", "dir", "amp", "delay")):
            assert_same(a, b, "%s excitation[%d] %s" % (what, c, n))
    uo, ur = o.upml_boxes(), r.upml_boxes()
    assert len(uo) == len(ur), "%s UPML boxes %d vs %d" % (what, len(uo), len(ur))
    for b, (x, y) in enumerate(zip(uo, ur)):
        assert x["start"] == y["start"] and x["n"] == y["n"], "%s UPML box %d geometry %s %s vs %s %s" % (what, b, x["start"], x["n"], y["start"], y["n"])
        for k in ("vv", "vvfn", "vvfo", "ii", "iifn", "iifo"):
            assert_same(x[k], y[k], "%s UPML box %d %s" % (what, b, k))
    mo, mr = o.mur_planes(), r.mur_planes()
    assert len(mo) == len(mr)
    for m, (x, y) in enumerate(zip(mo, mr)):
        for k in ("ny", "top", "line", "shift", "n", "start_ts"):
            assert x[k] == y[k], "%s Mur %d %s: %s vs %s" % (what, m, k, x[k], y[k])
        assert_same(x["coeff_nyP"], y["coeff_nyP"], "%s Mur %d nyP" % (what, m))
        assert_same(x["coeff_nyPP"], y["coeff_nyPP"], "%s Mur %d nyPP" % (what, m))
    lo, lr = o.lorentz(), r.lorentz()
    assert len(lo) == len(lr), "%s Lorentz order %d vs %d" % (what, len(lo), len(lr))
    for i, (x, y) in enumerate(zip(lo, lr)):
        assert x["count"] == y["count"] and x["flags"] == y["flags"], "%s Lorentz %d count/flags %s %s vs %s %s" % (what, i, x["count"], x["flags"], y["count"], y["flags"])
        assert_same(x["pos"], y["pos"], "%s Lorentz %d pos" % (what, i))
        for k in ("v_int", "v_ext", "v_lor", "i_int", "i_ext", "i_lor"):
            assert (x[k] is None) == (y[k] is None), "%s Lorentz %d %s presence" % (what, i, k)
            if x[k] is not None:
                assert_same(x[k], y[k], "%s Lorentz %d %s" % (what, i, k))
    to, tr = o.tfsf(), r.tfsf()
    if to is None:
        assert tr is None or not any(any(a) for a in tr["active"])
    else:
        assert tr is not None and to["start"] == tr["start"] and to["stop"] == tr["stop"] and to["active"] == tr["active"]
        assert to["max_delay"] == tr["max_delay"]
        assert sorted(to["faces"]) == sorted(tr["faces"])
        for k in to["faces"]:
            for a, b, n in zip(to["faces"][k], tr["faces"][k], ("delay", "delta", "amp")):
                assert_same(a, b, "%s TFSF %s %s" % (what, k, n))
    ao, ar = o.absorbing_sheets(), r.absorbing_sheets()
    assert len(ao) == len(ar)
    for i, (x, y) in enumerate(zip(ao, ar)):
        for k in ("ny", "type", "positive", "x0", "x1"):
            assert x[k] == y[k], "%s sheet %d %s" % (what, i, k)
        for k in ("K1P", "K1PP") + (("K2P", "K2PP") if x["type"] == 2 else ()):
            assert_same(x[k], y[k], "%s sheet %d %s" % (what, i, k))


def assert_state_equal(o, r, what=""):
    """E, H and the UPML flux after the same number of timesteps"""
    assert o.num_ts == r.num_ts
    assert_same(o.volt, r.volt, what + " volt @%d" % o.num_ts)
    assert_same(o.curr, r.curr, what + " curr @%d" % o.num_ts)
    for b in range(len(o.upml_boxes())):
        for c in (0, 1):
            assert_same(o.upml_flux(b, c), r.upml_flux(b, c), "%s UPML flux box %d %s" % (what, b, "curr" if c else "volt"))
