"""GPU parity tests: the CUDA engine, driven through the C ABI, against the CPU oracle on the
same operator.  Bar: bit-exact E/H fields (the reference's own cross-engine rule,
TESTSUITE/enginetests/cavity.m:155), which implies rel-L2 0 <= 1e-5 on every probe series."""
import numpy as np
import pytest

from oracle.pyoracle import BC_PEC, BC_PMC, BC_MUR, BC_PML
from tests import cases
from tests.gpu_util import operator_from_oracle, assert_fields_equal

pytestmark = pytest.mark.gpu


def run_both(s, steps=(1, 2, 17, 80), what="", **kw):
    """the oracle against ALL timestep schedules of the CUDA engine: one-pass requested (it runs where the hook set
    allows it), two-pass with one cell per thread (k_small_E/H, the automatic choice on meshes of this size) and
    two-pass with the float4 / z-march kernels (k_update_E/H, large meshes); returns the one-pass engine"""
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    eng.SetOption("fused", 1)
    eng2 = op.CreateEngine()
    eng2.SetOption("fused", 0)
    assert "fused_EH" not in [n for n, _ in eng2.TimeSchedule(0)] and eng2.GetOption("small") == 1
    eng3 = op.CreateEngine()
    eng3.SetOption("fused", 0)
    eng3.SetOption("small", 0)
    assert eng3.GetOption("small") == 0
    for k, v in kw.items():
        if k == "tuning":
            for e in (eng, eng2, eng3):
                e.SetTuning(*v)
    total = 0
    for n in steps:
        s.iterate(n)
        total += n
        for e, name in ((eng, "one-pass requested"), (eng2, "two-pass, cell per thread"), (eng3, "two-pass, float4 march")):
            e.IterateTS(n)
            assert e.GetNumberOfTimesteps() == s.num_ts == total
            mv, mc = assert_fields_equal(e, s, "%s (%s) after %d steps" % (what, name, total))
    assert mv > 0 and mc > 0, "fields stayed zero: the comparison would be vacuous"
    eng2.close()
    eng3.close()
    return eng


def test_stencil_pec_box():
    s = cases.uniform_box(n=(27, 11, 33), bc=(BC_PEC,) * 6)
    run_both(s, what="PEC box")


def test_engine_cavity_mur_pml_pmc():
    """the reference's engine test case: BC {MUR, PML_8, PMC, PEC, PEC, PEC}, dielectric box"""
    s = cases.engine_cavity()
    eng = run_both(s, steps=(1, 3, 50, 446), what="enginetests/cavity")
    for b, box in enumerate(s.upml_boxes()):
        for w in (0, 1):
            got = eng.GetUPMLFlux(b, w, box["n"])
            ref = s.upml_flux(b, w)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("n", [(23, 26, 21), (40, 36, 44), (130, 9, 7), (5, 5, 200)])
def test_all_pml_odd_sizes(n):
    pml = 4 if min(n) >= 12 else 1
    s = cases.uniform_box(n=n, bc=(BC_PML,) * 6, pml=pml)
    run_both(s, steps=(1, 5, 60), what="all-PML %s" % (n,))


def test_all_mur():
    s = cases.uniform_box(n=(31, 30, 29), bc=(BC_MUR,) * 6)
    run_both(s, steps=(1, 9, 120), what="all-Mur")


def test_no_graph_and_small_blocks():
    s = cases.engine_cavity()
    run_both(s, steps=(2, 41), what="tuned", tuning=(4, 5, 0))


def test_materials_and_metal():
    mats = [dict(start=(0.004, 0.003, 0.005), stop=(0.012, 0.010, 0.017), epsR=4.2, mueR=1.5, kappa=0.05, sigma=10.0)]
    metals = [dict(start=(0.015, 0.0, 0.010), stop=(0.020, 0.012, 0.010))]
    s = cases.uniform_box(n=(30, 24, 36), bc=(BC_PML, BC_PML, BC_MUR, BC_PEC, BC_PMC, BC_PML), pml=6,
                          materials=mats, metals=metals)
    run_both(s, steps=(1, 30, 150), what="materials")


def test_one_pass_and_two_pass_schedules_agree_and_are_used():
    """the one-pass timestep is the automatic choice from 20 M cells per GPU when the hook set allows it (smaller
    meshes run faster on the one-cell-per-thread two-pass kernels), the option forces either; both match the
    oracle bit for bit, also when toggled mid-run"""
    s = cases.engine_cavity()
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    assert eng.GetOption("fused") == 0 and eng.GetOption("small") == 1   # automatic choice on a small mesh
    big = cases.uniform_box(n=(200, 170, 130), bc=(BC_PML,) * 6, pml=8)
    eb = operator_from_oracle(big).CreateEngine()
    assert eb.GetOption("fused") == 0 and eb.GetOption("small") == 1     # 4.4 M cells: still two-pass, one cell per thread
    big.iterate(5)
    eb.IterateTS(5)
    assert_fields_equal(eb, big, "automatic two-pass (one cell per thread) on a 4.4 M cell mesh")
    eb.SetOption("fused_min_cells", 4000000)                              # the automatic rule with a lower threshold
    assert eb.GetOption("fused") == 1 and eb.GetOption("tma") == 1        # one-pass, TMA-staged
    big.iterate(7)
    eb.IterateTS(7)
    assert_fields_equal(eb, big, "automatic one-pass on a 4.4 M cell mesh")
    eb.SetOption("fused", 0)
    eb.SetOption("small", 0)                                              # the float4 / z-march two-pass kernels
    assert eb.GetOption("small") == 0
    big.iterate(4)
    eb.IterateTS(4)
    assert_fields_equal(eb, big, "float4 two-pass kernels on a 4.4 M cell mesh")
    eb.close()
    # requested: one-pass, with the UPML boxes on the two-pass shell around it
    eng.SetOption("fused", 1)
    names = [n for n, _ in eng.TimeSchedule(0)]
    assert "fused_EH" in names and "shell_E" in names and "shell_H" in names
    assert eng.GetOption("fused") == 1 and eng.GetOption("tma") == 1  # TMA-staged kernel, not the fallback
    s2 = cases.uniform_box(n=(27, 11, 33), bc=(BC_MUR, BC_MUR, BC_PMC, BC_PEC, BC_PEC, BC_MUR))
    e2 = operator_from_oracle(s2).CreateEngine()
    e2.SetOption("fused", 1)
    assert "fused_EH" in [n for n, _ in e2.TimeSchedule(0)]
    s2.iterate(90)
    e2.IterateTS(90)
    assert_fields_equal(e2, s2, "automatic one-pass, Mur/PMC/PEC")
    total = 0
    for n, fused, tma in ((3, 1, 1), (4, 0, 1), (5, 1, 0), (31, 1, 1), (2, 0, 0), (40, 1, 0), (7, 1, 1)):
        eng.SetOption("fused", fused)
        eng.SetOption("tma", tma)
        names = [x for x, _ in eng.TimeSchedule(0)]
        assert ("fused_EH" in names) == bool(fused)
        assert eng.GetOption("tma") == (1 if fused and tma else 0)
        s.iterate(n)
        eng.IterateTS(n)
        total += n
        assert_fields_equal(eng, s, "fused=%d after %d" % (fused, total))


def test_graded_mesh_wide_operator_index():
    """a graded mesh has (almost) one coefficient tuple per cell: more than 65535 tuples switch the
    per-cell operator index to 32 bit (all kernels have that instance, the TMA kernel runs a
    2-stage ring there); PML + Mur + PEC faces, both schedules, with and without TMA"""
    rng = np.random.default_rng(11)
    lines = [np.cumsum(1e-3 * (1.0 + 0.4 * rng.random(m))) for m in (50, 45, 41)]
    from oracle.pyoracle import OracleSim, EXC_E_SOFT
    s = OracleSim(lines[0], lines[1], lines[2], 1.0)
    s.set_bc([BC_PML, BC_MUR, BC_PML, BC_PEC, BC_PMC, BC_PML], (5,) * 6)
    s.set_excite_gauss(0.0, 299792458.0 / (20 * 1.4e-3))
    c = (float(lines[0][25]), float(lines[1][22]), float(0.5 * (lines[2][20] + lines[2][21])))
    s.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
    s.build()
    eng = run_both(s, steps=(1, 40, 120), what="graded mesh")
    assert eng.GetStats()["index_bytes"] == 4 and eng.GetOption("tma") == 1
    eng.SetOption("tma", 0)
    s.iterate(30)
    eng.IterateTS(30)
    assert_fields_equal(eng, s, "graded mesh, register-staged one-pass kernel")


def test_programmatic_dependent_launch():
    """option "pdl": every kernel of the timestep graph is launched with the programmatic-stream-serialization attribute
    and starts with griddepcontrol.wait; same bits as plain stream order, both schedules, mixed boundary conditions"""
    s = cases.engine_cavity()
    eng = operator_from_oracle(s).CreateEngine()
    try:
        for steps, pdl, fused in ((5, 1, 1), (20, 1, 0), (7, 0, 1), (40, 1, 1)):
            eng.SetOption("pdl", pdl)
            eng.SetOption("fused", fused)
            assert eng.GetOption("pdl") == pdl
            s.iterate(steps)
            eng.IterateTS(steps)
            assert_fields_equal(eng, s, "pdl=%d fused=%d" % (pdl, fused))
    finally:
        eng.SetOption("pdl", 0)   # process-wide switch


@pytest.mark.parametrize("n,pml", [((23, 26, 21), 4), ((150, 20, 18), 8), ((300, 12, 40), 8), ((12, 7, 60), 1)])
def test_x_slab_boxes_inside_the_one_pass_kernel(n, pml):
    """option "xslab": the UPML boxes at the x ends are updated by lanes of the TMA one-pass kernel
    (one cell per lane, flux of the voltages ping-ponged) instead of the shell launches; one, two
    and three x tiles; toggled mid-run at odd and even timestep counts; flux compared as well"""
    s = cases.uniform_box(n=n, bc=(BC_PML,) * 6, pml=pml)
    eng = operator_from_oracle(s).CreateEngine()
    eng.SetOption("fused", 1)
    eng.SetOption("xslab", 1)
    assert eng.GetOption("xslab") == 2 and eng.GetOption("tma") == 1
    total = 0
    for steps, xs in ((1, 1), (6, 1), (3, 0), (4, 1), (45, 1), (2, 0), (9, 1)):
        eng.SetOption("xslab", xs)
        assert eng.GetOption("xslab") == 2 * xs
        s.iterate(steps)
        eng.IterateTS(steps)
        total += steps
        assert_fields_equal(eng, s, "x slabs inline=%d after %d steps" % (xs, total))
        for b, box in enumerate(s.upml_boxes()):
            for w in (0, 1):
                assert np.array_equal(eng.GetUPMLFlux(b, w, box["n"]).view(np.uint32), s.upml_flux(b, w).view(np.uint32))


@pytest.mark.parametrize("n,pml,bc", [((150, 20, 18), 8, None), ((300, 33, 40), 8, None), ((64, 40, 33), 8, None), ((1031, 19, 21), 8, None),
                                      ((57, 37, 30), 4, None), ((96, 31, 34), 8, (BC_PML, BC_PML, BC_MUR, BC_PML, BC_PEC, BC_PML)),
                                      ((80, 30, 29), 6, (BC_PML, BC_PEC, BC_PML, BC_PML, BC_PML, BC_PMC))])
def test_x_slab_windows_tma_staged(n, pml, bc):
    """option "xslab" = 2: the UPML boxes at the x ends are updated, E and H in one pass, by the TMA-staged window
    kernel k_xslab_tma (16-line windows = whole 32-byte sectors, launched after the big kernel); window starts that
    are / are not multiples of 16 lines, partial last row tiles, mixed boundary conditions (Mur / PEC / PMC faces next
    to the slabs), toggled against the shell path at odd and even timestep counts, flux compared as well"""
    src = (n[0] - 22 if n[0] > 200 else n[0] // 2, n[1] // 2, n[2] // 2)
    s = cases.uniform_box(n=n, bc=bc or (BC_PML,) * 6, pml=pml, src_pos=src)
    eng = operator_from_oracle(s).CreateEngine()
    eng.SetOption("fused", 1)
    eng.SetOption("xslab", 2)
    assert eng.GetOption("tma") == 1
    nslabs = eng.GetOption("xslab")   # slabs with hook-changed cells inside (Mur faces) stay on the shell path
    assert nslabs == 2 if bc is None else nslabs in (0, 1, 2), nslabs
    total = 0
    for steps, xs in ((1, 2), (6, 2), (3, 0), (4, 2), (45, 2), (2, 1), (9, 2)):
        eng.SetOption("xslab", xs)
        s.iterate(steps)
        eng.IterateTS(steps)
        total += steps
        assert_fields_equal(eng, s, "x-slab windows mode %d after %d steps" % (xs, total))
        for b, box in enumerate(s.upml_boxes()):
            for w in (0, 1):
                assert np.array_equal(eng.GetUPMLFlux(b, w, box["n"]).view(np.uint32), s.upml_flux(b, w).view(np.uint32))
    names = [nm for nm, _ in eng.TimeSchedule(0)]
    assert ("xslab_EH" in names) == (nslabs > 0)


@pytest.mark.parametrize("case", ["line", "box"])
def test_local_absorbing_sheets(case):
    """Engine_Ext_Absorbing_BC (engine_ext_absorbing_bc.cpp:108-366) -- first order
    Mur sheets inside the mesh, with and without super-absorption, both normal signs, next to UPML and
    Mur faces; all six hooks in the reference's order, both schedules"""
    from oracle.pyoracle import OracleSim, EXC_E_SOFT
    if case == "line":
        # parallel-plate line along z (Rect_Waveguide_W_Local_Absorbers.py in small): sheets two lines from the ends
        x, y, z = np.arange(9) * 1e-3, np.arange(8) * 1e-3, np.arange(70) * 1e-3
        s = OracleSim(x, y, z, 1.0)
        s.set_bc([BC_PEC, BC_PEC, BC_PMC, BC_PMC, BC_PEC, BC_PEC])
        s.set_excite_gauss(8e9, 3e9)
        s.add_excitation((x[0], y[0], z[30]), (x[-1], y[-1], z[30]), EXC_E_SOFT, (1, 0, 0))
        s.add_absorbing_sheet((0, 0, 2), (8, 7, 2), True, 2, 0.0)
        s.add_absorbing_sheet((0, 0, 67), (8, 7, 67), False, 2, 3.2e8)
        steps = (1, 60, 400)
    else:
        x, y, z = np.arange(34) * 1e-3, np.cumsum(np.r_[0, np.linspace(1, 1.6, 27)]) * 1e-3, np.arange(30) * 1e-3
        s = OracleSim(x, y, z, 1.0)
        s.set_bc([BC_PML, BC_MUR, BC_PEC, BC_PML, BC_PMC, BC_PEC], (5,) * 6)
        s.set_excite_gauss(6e9, 4e9)
        c = cases.edge_center((x, y, z), 2, (15, 12, 14))
        s.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
        s.add_absorbing_sheet((8, 2, 20), (25, 2, 27), True, 1, 0.0)        # y-normal, plain Mur
        s.add_absorbing_sheet((3, 3, 27), (30, 20, 27), False, 2, 0.0)      # z-normal, super-absorbing, reaches into the UPML
        s.add_absorbing_sheet((29, 4, 3), (29, 18, 12), False, 2, 2.5e8)    # x-normal
        steps = (1, 25, 120)
    s.build()
    assert len(s.absorbing_sheets()) in (2, 3)
    eng = run_both(s, steps=steps, what="absorbing sheets " + case)
    names = [n for n, _ in eng.TimeSchedule(0)]
    assert "fused_EH" in names and "sheet_apply_V" in names and "sheet_apply_I" in names


@pytest.mark.parametrize("prop_dir,e_amp,periodic", [((1, 0.3, 0.2), (0, 0, 1), False), ((0, 0, -1), (1, 1, 0), False), ((-0.4, 1, 0), (0, 0, 2), True)])
def test_tfsf_plane_wave(prop_dir, e_amp, periodic):
    """Engine_Ext_TFSF (engine_ext_tfsf.cpp:36-215, tables of operator_ext_tfsf.cpp:86-406)
    -- plane-wave injection on the six faces of a box inside an all-UPML mesh: oblique, axial and
    periodic (sinusoidal) cases; shared box edges get their two updates in the reference's order"""
    from oracle.pyoracle import OracleSim
    x, y, z = np.arange(38) * 1e-3, np.cumsum(np.r_[0, np.linspace(1, 1.3, 33)]) * 1e-3, np.arange(36) * 1e-3
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_PML] * 6, (6,) * 6)
    if periodic:
        s.set_excite_sinus(6e9)
    else:
        s.set_excite_gauss(5e9, 4e9)
    s.set_tfsf((10, 9, 10), (27, 24, 25), prop_dir, e_amp)
    s.build()
    t = s.tfsf()
    assert t is not None and t["max_delay"] > 5 and len(t["faces"]) == 24
    eng = run_both(s, steps=(1, 7, 60, 150), what="TFSF %s" % (prop_dir,))
    names = [n for n, _ in eng.TimeSchedule(0)]
    assert "fused_EH" in names and "tfsf_V" in names and "tfsf_I" in names
    v = s.volt
    inside = np.abs(v[:, 13:25, 12:22, 13:23]).max()
    outside = max(np.abs(v[:, 7:9]).max(), np.abs(v[:, 30:32]).max())
    assert inside > 1e-4 and outside < 0.02 * inside  # total field inside, (almost) nothing scattered outside


from tests.golden.make_golden import CASES as GOLDEN_CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
@pytest.mark.parametrize("fused", [0, 1])
def test_engine_reproduces_committed_golden_vectors(name, fused):
    """the CUDA engine against the committed fixtures of tests/golden, which were produced by the REFERENCE'S OWN
    multithreaded engine (oracle/_ref, see make_golden.py): probe series bit for bit, field digests at three
    timesteps, both schedules.  The operator tables are built by the oracle (pinned to the reference's builders by
    tests/test_ref_pinning.py) and uploaded through the C ABI."""
    import os
    from tests.golden import make_golden as G
    g = np.load(os.path.join(os.path.dirname(G.__file__), name + ".npz"))
    s = GOLDEN_CASES[name][0]()
    assert s.dT == float(g["dT"])
    eng = operator_from_oracle(s).CreateEngine()
    try:
        eng.SetOption("fused", fused)
    except Exception:
        pytest.skip("schedule not available for this hook set")
    if eng.GetOption("fused") != fused:
        pytest.skip("schedule not available for this hook set")
    probes = [tuple(map(tuple, p)) for p in g["probes"].tolist()]
    ids = [eng.AddVoltageProbe(a, b) for a, b in probes]
    steps = int(g["steps"])
    want = {int(t): (int(dv), int(di)) for t, dv, di in g["digests"].tolist()}
    series = np.zeros((steps, len(probes)))
    for t in range(steps):
        eng.IterateTS(1)
        u = eng.ReadProbes()
        series[t] = [u[i] for i in ids]
        if t + 1 in want:
            assert (int(G.field_digest(eng.GetFields(0))), int(G.field_digest(eng.GetFields(1)))) == want[t + 1], "fields at step %d" % (t + 1)
    assert np.array_equal(series.view(np.uint64), g["series"].view(np.uint64))
