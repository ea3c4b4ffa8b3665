"""The product's host-side operator builder (csrc/host/synthetic_operator.cpp)
against (a) the REFERENCE'S OWN Operator::CalcECOperator + extension builders (oracle/_ref, the unmodified
reference translation units) and (b) the oracle's dense restatement: timestep, every coefficient, UPML aux
coefficients, excitation lists and Mur coefficients, bit for bit.
CPU only (host code of libopenems_b200.so; no kernel is launched)."""
import numpy as np
import pytest

from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT, EXC_E_HARD, EXC_H_SOFT
from oracle import pyref
from openems_b200 import SyntheticOperator

AGAINST = [pytest.param("reference", marks=pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not available")), "oracle"]


@pytest.fixture(params=AGAINST)
def against(request):
    return request.param


def both(lines, setup, unit=1.0, against="oracle"):
    o = pyref.RefSim(*lines, unit) if against == "reference" else OracleSim(*lines, unit)
    p = SyntheticOperator(*lines, unit)
    setup(o)
    setup(p)
    o.build(10 ** 6)
    p.build(10 ** 6)
    return o, p


def compare(o, p):
    assert p.dT == o.dT
    assert p.nyquist == o.nyquist
    for w in ("vv", "vi", "ii", "iv"):
        a, b = p.dense(w), o.coeff(w)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), w
    sv, si, _ = o.signal()
    pv, pi = p.signal()
    assert np.array_equal(sv, pv) and np.array_equal(si, pi)
    for w in (0, 1):
        for a, b in zip(o.excitation(w), p.excitation(w)):
            assert np.array_equal(a, b)
    om, pm = o.mur_planes(), p.mur_planes()
    assert len(om) == len(pm)
    for a, b in zip(om, pm):
        assert (a["ny"], a["line"], a["shift"], a["start_ts"]) == (b["ny"], b["line"], b["shift"], b["start_ts"])
        assert np.array_equal(a["coeff_nyP"], b["coeff_nyP"]) and np.array_equal(a["coeff_nyPP"], b["coeff_nyPP"])
    ob, pb = o.upml_boxes(), p.upml_boxes()
    assert [(b["start"], b["n"]) for b in ob] == [(b["start"], b["n"]) for b in pb]
    pml = p.dense("pml")
    inbox = np.zeros(pml.shape, bool)
    names = dict(vv="pml_vv", vvfn="pml_vvfn", vvfo="pml_vvfo", ii="pml_ii", iifn="pml_iifn", iifo="pml_iifo")
    dense_aux = {k: p.dense(v) for k, v in names.items()}
    for b in ob:
        s, n = b["start"], b["n"]
        sl = (slice(s[0], s[0] + n[0]), slice(s[1], s[1] + n[1]), slice(s[2], s[2] + n[2]))
        inbox[sl] = True
        for k in names:
            got = dense_aux[k][(slice(None),) + sl]
            assert np.array_equal(got.view(np.uint32), b[k].view(np.uint32)), k
    assert np.array_equal(pml != 0, inbox)


def test_vacuum_all_pml(against):
    # mesh in drawing units (mm) with unit 1e-3, like the tutorials: exact, equal spacings
    lines = tuple(np.arange(n, dtype=np.float64) for n in (30, 28, 40))

    def setup(s):
        s.set_bc([BC_PML] * 6, (8,) * 6)
        s.set_excite_gauss(0.0, 15e9)
        s.add_excitation((14.5, 14, 16), (14.5, 14, 16), EXC_E_SOFT, (1, 0, 0))
    o, p = both(lines, setup, unit=1e-3, against=against)
    compare(o, p)
    assert p.n_unique < 6000 and p.index_bytes == 2  # ~17^3 depth classes
    assert p.unique_planes < 40  # interior planes away from the source share one computation
    # the plane representation the engine receives (oems_cuda_set_operator_planes) is the same index
    up, ids = p.planes()
    assert up.shape == (p.unique_planes, 28, 30) and ids.shape == (40,) and ids.max() < p.unique_planes
    assert np.array_equal(up[ids], p.index())


def test_mixed_bc_materials_metal_nonuniform_mesh(against):
    x = np.cumsum(np.r_[0, np.full(10, 1.0), np.linspace(1.0, 0.5, 6), np.full(12, 0.5)]) * 1e-3
    y = np.arange(26) * 0.8e-3
    z = np.cumsum(np.r_[0, np.full(30, 0.7)]) * 1e-3

    def setup(s):
        s.set_bc([BC_MUR, BC_PML, BC_PMC, BC_PEC, BC_PML, BC_MUR], (8, 6, 8, 8, 5, 8))
        s.set_background(1.0, 1.0, 0.0, 0.0)
        s.set_excite_gauss(4e9, 3e9)
        s.add_material((x[4], y[3], z[6]), (x[15], y[12], z[15]), epsR=3.66, mueR=1.2, kappa=0.01, sigma=5.0)
        s.add_material((x[10], y[8], z[10]), (x[20], y[20], z[22]), epsR=2.0, prio=2)
        s.add_metal((x[6], y[5], z[18]), (x[18], y[15], z[18]))
        s.add_excitation((x[12], y[10], z[3]), (x[12], y[16], z[3]), EXC_E_HARD, (0, 1, 0), delay=2e-11)
        s.add_excitation((x[3], y[3], z[8]), (x[5], y[5], z[8]), EXC_H_SOFT, (1, 1, 0))
    o, p = both((x, y, z), setup, against=against)
    compare(o, p)


def test_excitation_on_mur_plane_delays_start(against):
    lines = tuple(np.arange(n) * 1e-3 for n in (14, 15, 20))

    def setup(s):
        s.set_bc([BC_PEC, BC_PEC, BC_PEC, BC_PEC, BC_MUR, BC_MUR])
        s.set_excite_gauss(10e9, 8e9)
        s.add_excitation((0, 0, 0), (0.013, 0.0135, 0), EXC_E_SOFT, (0, 1, 0))
    o, p = both(lines, setup, against=against)
    compare(o, p)
    assert p.mur_planes()[0]["start_ts"] > 0


def test_lorentz_lists_match(against):
    lines = tuple(np.arange(n) * 1e-3 for n in (20, 22, 24))

    def setup(s):
        s.set_bc([BC_PML] * 6, (4,) * 6)
        s.set_excite_gauss(5e9, 5e9)
        s.add_lorentz((0.006, 0.007, 0.008), (0.013, 0.014, 0.015), epsR=1.0, eps_fp=(5e9, 2e9), eps_tau=(5e-9, 0.0),
                      eps_flor=(0.0, 7e9), mue_fp=(5e9,), mue_tau=(5e-9,))
        s.add_excitation((0.003, 0.003, 0.0035), (0.003, 0.003, 0.0035), EXC_E_SOFT, (0, 0, 1))
    o, p = both(lines, setup, against=against)
    compare(o, p)
    ol = o.lorentz()
    assert p.lorentz_counts() == [L["count"] for L in ol]
    assert ol[0]["count"] > 0


def test_large_mesh_builds_fast_without_dense_arrays():
    import time
    lines = tuple(np.arange(n, dtype=np.float64) for n in (256, 256, 256))
    p = SyntheticOperator(*lines, 1e-3)
    p.set_bc([BC_PML] * 6)
    p.set_excite_gauss(0.0, 15e9)
    p.add_excitation((127.5, 128, 128), (127.5, 128, 128), EXC_E_SOFT, (1, 0, 0))
    t0 = time.time()
    p.build()
    dt = time.time() - t0
    assert p.unique_planes <= 2 * 12 + 8
    assert p.index().shape == (256, 256, 256)
    assert dt < 60


def test_slab_restricted_build_equals_the_full_build():
    """one process per GPU: each rank builds only the planes it holds; the timestep is agreed on by a MIN-reduction of
    the per-rank values; tuples, plane contents and ids of the held planes equal the full build"""
    lines = tuple(np.arange(n, dtype=np.float64) for n in (30, 28, 48))

    def setup(s):
        s.set_bc([BC_PML] * 6, (8,) * 6)
        s.set_excite_gauss(0.0, 15e9)
        s.add_material((5, 5, 30), (20, 20, 40), epsR=2.0)
        s.add_excitation((14.5, 14, 16), (14.5, 14, 16), EXC_E_SOFT, (1, 0, 0))
    full = SyntheticOperator(*lines, 1e-3)
    setup(full)
    full.build()
    tab_full = full.table()
    idx_full = full.index()
    bounds = [0, 14, 30, 48]
    parts = []
    for r in range(3):
        p = SyntheticOperator(*lines, 1e-3)
        setup(p)
        p.set_slab(bounds[r], bounds[r + 1])
        parts.append(p)
    dT = min(p.local_timestep() for p in parts)
    assert dT == full.dT
    for r, p in enumerate(parts):
        p.set_timestep(dT)
        p.build()
        assert p.dT == full.dT and p.unique_planes <= full.unique_planes
        up, ids = p.planes()
        tab = p.table()
        lo, hi = max(bounds[r] - 1, 0), min(bounds[r + 1] + 1, 48)
        for k in range(lo, hi):
            a = tab[up[ids[k]].astype(np.int64)]
            b = tab_full[idx_full[k].astype(np.int64)]
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (r, k)
    assert sum(p.unique_planes for p in parts) >= full.unique_planes
