"""Pins the CPU oracle (oracle/fdtd_oracle.c) to the REFERENCE ITSELF.

oracle/_ref/libopenems_ref.so is built from the unmodified translation units of /root/reference
(oracle/Makefile.ref): Operator::CalcECOperator, all operator-extension builders, Engine / Engine_sse /
Engine_SSE_Compressed / Engine_Multithread, all engine extensions and the Processing classes.  Every test
pushes one case through the restatement and through that library and demands equality of EVERY BIT:
timestep, vv/vi/ii/iv, excitation signal and lists, UPML / Mur / Lorentz / TFSF / sheet tables, E, H, UPML
flux after N steps, and the read-outs.  (TESTSUITE/enginetests/cavity.m:155 is the reference's own rule for
engine variants; here it is applied between the reference's engines and the restatement.)
CPU only; runs wherever the prebuilt library or the reference tree is present."""
import os

import numpy as np
import pytest

from oracle import pyref
from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT, EXC_E_HARD, EXC_H_SOFT, EXC_H_HARD
from oracle.pyref import RefSim, ENGINE_BASIC, ENGINE_SSE, ENGINE_SSE_COMPRESSED, ENGINE_MULTITHREADED
from tests import cases, configs
from tests.ref_util import build_both, assert_operator_equal, assert_state_equal, assert_same, backend, ref_class

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built and no reference tree")

C0 = 299792458.0
ENGINES = [ENGINE_BASIC, ENGINE_SSE, ENGINE_SSE_COMPRESSED, ENGINE_MULTITHREADED]
ENGINE_IDS = ["basic", "sse", "sse-compressed", "multithreaded"]


def _sims(res):
    return (res[0][0], res[1][0]) if isinstance(res[0], tuple) else res


def _run(fn, *a, steps=(1, 49), engine=ENGINE_BASIC, threads=3, **kw):
    o, r = _sims(build_both(fn, *a, engine=engine, threads=threads, **kw))
    assert_operator_equal(o, r)
    for n in steps:
        o.iterate(n)
        r.iterate(n)
        assert_state_equal(o, r)
    assert np.abs(o.volt).max() > 0 and np.abs(o.curr).max() > 0
    return o, r


def test_reference_library_is_the_unmodified_reference():
    assert b"unmodified" in pyref.lib().ref_version()


@pytest.mark.parametrize("engine", ENGINES, ids=ENGINE_IDS)
def test_enginetest_cavity_all_reference_engines(engine):
    """TESTSUITE/enginetests/cavity.m:70-115: {MUR, PML_8, PMC, PEC, PEC, PEC} + dielectric box"""
    _run(cases.engine_cavity, steps=(1, 2, 197), engine=engine)


@pytest.mark.parametrize("engine", ENGINES, ids=ENGINE_IDS)
def test_all_pml_odd_sizes(engine):
    _run(cases.uniform_box, n=(29, 23, 31), steps=(1, 59), engine=engine)


def test_analytic_cavity_case_operator_and_fields():
    o, r = _run(cases.analytic_cavity, 2000, steps=(300,))
    for a, b in (((5, 4, 5), (6, 4, 5)), ((12, 4, 5), (12, 4, 6))):
        assert o.voltage_integral(a, b) == r.voltage_integral(a, b)


@pytest.mark.parametrize("excite", ["gauss", "sinus"])
def test_c1_parallel_plate_waveguide(excite):
    """Mur with delayed start (the source plane touches the z-Mur planes' neighbourhood), PMC, periodic signal"""
    _run(configs.c1_parallel_plate_waveguide, excite, steps=(1, 120), engine=ENGINE_MULTITHREADED)


def test_c2_msl_notch_filter():
    _run(configs.c2_msl_notch_filter, steps=(40,), engine=ENGINE_MULTITHREADED)


def test_c3_patch_antenna_lumped_rc():
    _run(configs.c3_patch_antenna, steps=(40,), engine=ENGINE_SSE_COMPRESSED)


@pytest.mark.parametrize("engine", [ENGINE_BASIC, ENGINE_MULTITHREADED], ids=["basic", "multithreaded"])
def test_c4_drude_block(engine):
    _run(configs.c4_drude_block, n=(30, 30, 30), block=(10, 20), steps=(1, 60), engine=engine)


def test_lorentz_two_pole_and_drude():
    n = (34, 30, 38)
    lor = [dict(start=(0.010, 0.008, 0.012), stop=(0.022, 0.020, 0.026), eps_fp=(5e9,), eps_tau=(5e-9,),
                mue_fp=(5e9,), mue_tau=(5e-9,)),
           dict(start=(0.004, 0.004, 0.004), stop=(0.008, 0.012, 0.010), epsR=2.0, eps_fp=(3e9, 6e9), eps_tau=(2e-9, 0.0),
                eps_flor=(0.0, 9e9), prio=3)]
    fc = C0 / (20 * 1e-3) / 2
    o, r = _run(cases.uniform_box, n=n, bc=(BC_PML,) * 6, pml=6, f0=fc, fc=fc, lorentz=lor, src_pos=(6, 15, 19), steps=(1, 40, 100))
    assert len(o.lorentz()) == 2


def test_mixed_bc_pml_sizes_materials_metal():
    """mixed PML sizes (upper-side grading quirk, operator_ext_upml.cpp:319), lossy dielectric, metal sheet"""
    lines = (np.arange(40, dtype=np.float64), np.arange(33, dtype=np.float64), np.arange(45, dtype=np.float64))

    def case():
        q = cases.OracleSim(*lines, 1e-3)
        q.set_bc([BC_PML, BC_PML, BC_MUR, BC_PML, BC_PMC, BC_PML], (8, 8, 8, 6, 8, 7))
        q.set_excite_gauss(6e9, 6e9)
        q.add_material((10, 5, 8), (25, 20, 30), epsR=2.5, kappa=0.01)
        q.add_material((14, 8, 10), (20, 12, 20), epsR=1.0, mueR=2.0, sigma=30.0, prio=2)
        q.add_metal((12, 10, 20), (30, 18, 20))
        q.add_excitation((20, 16, 10.5), (20, 16, 10.5), EXC_E_SOFT, (0, 0, 1))
        q.build()
        return q
    _run(case, steps=(1, 30, 120))


def test_graded_mesh_background_material_timestep_factor_mur_velocity():
    rng = np.random.default_rng(3)
    x = np.cumsum(np.r_[0, 1 + 0.5 * rng.random(25)]) * 1e-3
    y = np.cumsum(np.r_[0, 1 + 0.3 * rng.random(19)]) * 1e-3
    z = np.cumsum(np.r_[0, np.linspace(2, 1, 29)]) * 1e-3

    def case():
        q = cases.OracleSim(x, y, z, 1.0)
        q.set_bc([BC_MUR, BC_MUR, BC_PEC, BC_PMC, BC_MUR, BC_PML], (8,) * 6)
        q.set_background(2.2, 1.0, 1e-3, 0.0)
        q.set_timestep(0.0, 0.7)
        q.set_mur_phase_velocity(C0 / 1.6)
        q.set_excite_gauss(4e9, 3e9)
        c = cases.edge_center((x, y, z), 2, (12, 9, 14))
        q.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
        q.build()
        return q
    _run(case, steps=(1, 80))


@pytest.mark.parametrize("exc_type", [EXC_E_SOFT, EXC_E_HARD, EXC_H_SOFT, EXC_H_HARD])
def test_excitation_types_and_delay(exc_type):
    """soft/hard E and H box sources with a delay; dirac and step signals are covered below"""
    def case():
        lines = tuple(np.arange(m) * 1e-3 for m in (22, 20, 24))
        q = cases.OracleSim(*lines, 1.0)
        q.set_bc([BC_PEC, BC_PMC, BC_MUR, BC_MUR, BC_PML, BC_PML], (5,) * 6)
        q.set_excite_gauss(5e9, 5e9)
        q.add_excitation((0.008, 0.007, 0.010), (0.012, 0.011, 0.010), exc_type, (1, 0.5, 0), delay=17e-12)
        q.add_excitation((0.005, 0.005, 0.015), (0.005, 0.009, 0.015), EXC_E_SOFT, (0, 1, 0), prio=1)
        q.build()
        return q
    _run(case, steps=(1, 70))


@pytest.mark.parametrize("signal", ["dirac", "step", "sinus"])
def test_signal_types(signal):
    def case():
        lines = tuple(np.arange(m) * 1e-3 for m in (18, 17, 19))
        q = cases.OracleSim(*lines, 1.0)
        q.set_bc([BC_MUR] * 6)
        getattr(q, "set_excite_" + signal)(8e9)
        c = cases.edge_center(lines, 2, (8, 8, 9))
        q.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
        q.build()
        return q
    _run(case, steps=(1, 50))


def test_lumped_rlc_raw_coefficients():
    rng = np.random.default_rng(7)
    cnt = 5
    pos = np.array([[8, 9, 10, 11, 12], [10, 10, 11, 11, 12], [12, 13, 12, 13, 14]], np.uint32)
    d = np.array([0, 1, 2, 2, 1], np.int32)
    co = {k: (rng.uniform(-0.3, 0.3, cnt)).astype(np.float32) for k in ("ilv", "i2v", "vv2", "vj1", "vj2", "ib0", "b1", "b2")}
    co["vvd"] = rng.uniform(0.5, 1.0, cnt).astype(np.float32)
    _run(cases.uniform_box, n=(24, 22, 26), bc=(BC_MUR,) * 6, extra=lambda s, lines: s.add_rlc_raw(d, pos, co),
         src_pos=(9, 10, 12), steps=(1, 2, 3, 50))


def test_steady_state_detection():
    s0, _ = configs.c1_parallel_plate_waveguide("sinus")
    period = s0.signal()[2]
    N = s0.N
    pos3, dirs = [], []
    for p in ((N[0] // 2, N[1] // 2, N[2] // 2), (0, N[1] // 2, N[2] // 2), (N[0] // 2, 0, N[2] // 2)):
        for n in range(3):
            pos3.append(p)
            dirs.append(n)
    pos3 = np.array(pos3, np.uint32).T.copy()
    dirs = np.array(dirs, np.int32)

    def case():
        q = configs.OracleSim(s0.x, s0.y, s0.z, 1.0)
        q.set_bc([BC_PMC, BC_PMC, BC_PEC, BC_PEC, BC_MUR, BC_MUR])
        q.set_excite_sinus(10e6)
        q.add_excitation((-10, -10, 0), (10, 10, 0), EXC_E_SOFT, (0, 1, 0))
        q.add_steadystate(period, pos3, dirs)
        q.build()
        return q
    o = case()
    # the scalar engine interface: its CalcFastEnergy (engine_interface_fdtd.cpp:302-347) is what the oracle restates;
    # the sse interface sums in four float lanes, which moves the energy-ratio part of the criterion by ~1 %
    with backend(ref_class(ENGINE_BASIC)):
        r = case()
    assert_operator_equal(o, r)
    seen = 0
    for it in range(8):
        n = period if it else period + 1
        o.iterate(n)
        r.iterate(n)
        a, b = o.steadystate_last_diff(), r.steadystate_last_diff()
        assert a == pytest.approx(b, rel=1e-12, abs=1e-300), (it, a, b)
        seen += b > 0
    assert seen >= 5
    assert_state_equal(o, r)


@pytest.mark.parametrize("prop,amp", [((0.0, 0.0, 1.0), (1.0, 0.0, 0.0)), ((1.0, 1.0, 0.5), (1.0, -1.0, 0.0))])
def test_tfsf_plane_wave(prop, amp):
    def extra(s, lines):
        s.set_tfsf((8, 8, 8), (21, 19, 23), prop, amp)
    # the point source of uniform_box sits outside the TFSF box (inside, the type-10 excitation box would shadow it)
    o, r = _run(cases.uniform_box, n=(30, 28, 32), bc=(BC_PML,) * 6, pml=6, extra=extra, src_pos=(7, 7, 7), steps=(1, 90))
    assert o.tfsf() is not None


@pytest.mark.parametrize("abc_type", [1, 2])
def test_local_absorbing_sheets(abc_type):
    def extra(s, lines):
        s.add_absorbing_sheet((4, 4, 5), (19, 17, 5), True, abc_type)
        s.add_absorbing_sheet((4, 4, 24), (19, 17, 24), False, abc_type, C0 / 1.2)
        s.add_absorbing_sheet((6, 3, 8), (6, 18, 20), False, abc_type)
    o, r = _run(cases.uniform_box, n=(24, 22, 30), bc=(BC_PEC,) * 6, extra=extra, steps=(1, 2, 100))
    assert len(o.absorbing_sheets()) == 3


def test_mur_delayed_start():
    """an excitation ON a Mur plane delays that plane (engine_ext_mur_abc.cpp:44-60)"""
    def case():
        lines = tuple(np.arange(m) * 1e-3 for m in (20, 18, 22))
        q = cases.OracleSim(*lines, 1.0)
        q.set_bc([BC_MUR] * 6)
        q.set_excite_gauss(6e9, 6e9)
        q.add_excitation((0.0, 0.004, 0.005), (0.0, 0.012, 0.015), EXC_E_SOFT, (0, 1, 0))   # on the xmin plane
        q.build()
        return q
    o, r = _run(case, steps=(1, 30))
    starts = [m["start_ts"] for m in r.mur_planes()]
    assert starts[0] > 0 and starts[1] == 0
    sig_len = len(r.signal()[0])
    o.iterate(starts[0] + 5 - o.num_ts)
    r.iterate(starts[0] + 5 - r.num_ts)
    assert_state_equal(o, r)
    assert o.num_ts > sig_len


def test_readouts_through_reference_engine_interface():
    """Engine_Interface_FDTD::CalcVoltageIntegral / GetRawField / CalcFastEnergy, ProcessCurrent::CalcIntegral and
    ProcessFields::CalcField (3 interpolations, E and H) against the restated read-outs"""
    o, r = _run(cases.engine_cavity, steps=(150,), engine=ENGINE_SSE_COMPRESSED)
    for a, b in (((5, 4, 5), (9, 4, 5)), ((5, 8, 5), (5, 3, 5)), ((12, 4, 5), (12, 4, 17))):
        assert o.voltage_integral(a, b) == r.voltage_integral(a, b) != 0
    for (a, b), nd, si, ei in ((((4, 3, 10), (12, 8, 10)), 2, (1, 1, 1), (1, 1, 1)), (((8, 2, 6), (8, 8, 20)), 0, (1, 1, 1), (1, 0, 1)),
                               (((3, 5, 4), (20, 5, 25)), 1, (1, 1, 0), (1, 1, 1))):
        assert o.current_integral(a, b, nd, si, ei) == r.current_integral(a, b, nd, si, ei) != 0
    for p in ((16, 6, 20), (0, 0, 0), (26, 10, 32), (25, 9, 31)):
        for h in (0, 1):
            assert np.array_equal(o.raw_field(h, p), r.raw_field(h, p))
    eo, er = o.energy(), r.energy()
    assert er > 0 and abs(eo - er) <= 1e-4 * er   # the sse interface sums in 4 float lanes (engine_interface_sse_fdtd.cpp:40-75)
    with backend(ref_class(ENGINE_BASIC)):
        rb = cases.engine_cavity()
    rb.iterate(150)
    assert o.energy() == rb.energy()              # the scalar interface (engine_interface_fdtd.cpp:302-347) is restated exactly
    start, stop = (0, 0, 0), tuple(n - 1 for n in o.N)
    for h in (0, 1):
        for interp in (0, 1, 2):
            assert_same(o.dump_field(h, interp, start, stop), r.dump_field(h, interp, start, stop), "dump H=%d interp=%d" % (h, interp))
    sub = ((3, 2, 4), (20, 8, 29))
    assert_same(o.dump_field(0, 2, *sub), r.dump_field(0, 2, *sub), "sub-box dump")


def test_multithreaded_engine_thread_counts():
    """Engine_Multithread with 1, 2, 5 threads == the scalar restatement (x-slab partition, useful.cpp:45-75)"""
    for th in (1, 2, 5):
        _run(cases.engine_cavity, steps=(60,), engine=ENGINE_MULTITHREADED, threads=th)


def test_sse_compressed_operator_dedup_count():
    """Operator_SSE_Compressed::CompressOperator (a3): tuple count of the reference == restated sse-compressed engine"""
    from oracle.pyoracle import OracleSSE
    o, r = _sims(build_both(cases.engine_cavity, engine=ENGINE_SSE_COMPRESSED))
    e = OracleSSE(o, 1)
    assert e.unique == r.sse_unique > 10
    e.iterate(77)
    r.iterate(77)
    v, c = e.fields()
    assert_same(v, r.volt, "sse restatement volt")
    assert_same(c, r.curr, "sse restatement curr")


def test_reference_processing_classes_probe_files(tmp_path):
    """the reference's own ProcessVoltage / ProcessCurrent / ProcessFieldProbe run by the RunFDTD loop write the
    series the restated integrals produce at the same timesteps (12 significant digits, processing.cpp:34)"""
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with backend(ref_class(ENGINE_MULTITHREADED, 2)):
            r = cases.engine_cavity()
        o = cases.engine_cavity()
        lines = (o.x, o.y, o.z)
        vbox = ((5, 4, 5), (9, 4, 5))
        cbox = ((4, 3, 10), (12, 8, 10))
        fpos = (16, 6, 20)
        r.add_probe(0, "ut1", [lines[n][vbox[0][n]] for n in range(3)], [lines[n][vbox[1][n]] for n in range(3)])
        # current box given on the dual mesh: snap to the same dual indices by using the dual line coordinates
        c0 = [o.disc_line(n, cbox[0][n], True) for n in range(3)]
        c1 = [o.disc_line(n, cbox[1][n], True) for n in range(3)]
        r.add_probe(1, "it1", c0, c1, norm_dir=2)
        fp = [lines[n][fpos[n]] for n in range(3)]
        r.add_probe(2, "et1", fp, fp)
        nr = 200
        r.run(nr)
        ut = pyref.read_probe_file("ut1")
        it = pyref.read_probe_file("it1")
        et = pyref.read_probe_file("et1")
        interval = max(1, r.nyquist // 4)
        assert len(ut) >= nr // interval
        rows_u, rows_i, rows_e = [], [], []
        # Processing::Process is first called at TS 0 (openems.cpp:1424)
        ts = 0
        while ts <= nr:
            if ts > 0:
                o.iterate(ts - o.num_ts)
            rows_u.append((ts * o.dT, o.voltage_integral(*vbox)))
            rows_i.append(((ts + 0.5) * o.dT, o.current_integral(cbox[0], cbox[1], 2)))
            rows_e.append((ts * o.dT,) + tuple(o.raw_field(0, fpos)))
            ts += interval
        k = min(len(rows_u), len(ut))
        assert k >= nr // interval
        def close(a, b):
            return np.allclose(a, b, rtol=1e-11, atol=1e-300)   # 12 significant digits in the file
        assert close(ut[:k, 0], [x[0] for x in rows_u[:k]]) and close(ut[:k, 1], [x[1] for x in rows_u[:k]])
        assert close(it[:k, 0], [x[0] for x in rows_i[:k]]) and close(it[:k, 1], [x[1] for x in rows_i[:k]])
        assert close(et[:k, 1:4], np.array([x[1:] for x in rows_e[:k]]))
        assert np.abs(ut[:, 1]).max() > 0 and np.abs(it[:, 1]).max() > 0
    finally:
        os.chdir(cwd)
