"""Pins the CPU oracle (oracle/) against what the reference's own tests hold for this path
(SURVEY.md 8c): the analytic cavity resonances, the cross-engine bit-equality rule, the
probe==dump rule and the uniform-mesh closed forms.  CPU only."""
import numpy as np
import pytest

from oracle.pyoracle import OracleSim, OracleSSE, BC_PEC, BC_PMC, BC_MUR, BC_PML
from tests import cases

C0 = 299792458.0
EPS0 = 8.85418781762e-12
MUE0 = 1.256637062e-6


def fft_time2freq(t, val):
    """matlab/FFT_time2freq.m"""
    dt = t[1] - t[0]
    L = len(val)
    nfft = 2 ** int(np.ceil(np.log2(L)))
    V = np.fft.fft(val, nfft) * dt
    f = 1 / (2 * dt) * np.linspace(0, 1, nfft // 2 + 1)
    return f, 2 * V[: nfft // 2 + 1]


def check_frequency(f, val, f_upper, f_lower, rel_amplitude, kind):
    """TESTSUITE/helperscripts/check_frequency.m"""
    max1 = val.max()
    nearest = lambda x: int(np.argmin(np.abs(f - x)))
    for f1, f2 in zip(f_lower, f_upper):
        seg = val[nearest(f1): nearest(f2) + 1]
        if kind == "inside" and seg.max() < max1 * rel_amplitude:
            return False
        if kind == "outside" and seg.max() > max1 * rel_amplitude:
            return False
    return True


@pytest.mark.slow
def test_analytic_cavity_resonances():
    """TESTSUITE/combinedtests/cavity.m:24-32,131-232"""
    s, probes, (a, b, d) = cases.analytic_cavity(20000)
    step = s.nyquist // 4  # openems.cpp:568, OverSampling 4 (:119)
    t, uy, uz = [], [], []
    n = 0
    while n < 20000:
        s.iterate(step)
        n += step
        t.append(n * s.dT)
        uy.append(s.voltage_integral(*probes["ut1y"]))
        uz.append(s.voltage_integral(*probes["ut1z"]))
    t = np.array(t)
    i0 = int(np.argmin(np.abs(t - 7e-10)))
    f, UY = fft_time2freq(t[i0:], np.array(uy)[i0:])
    _, UZ = fft_time2freq(t[i0:], np.array(uz)[i0:])
    f_start, f_stop = 1e9, 10e9
    i1, i2 = int(np.argmin(np.abs(f - f_start))), int(np.argmin(np.abs(f - f_stop)))
    f, UY, UZ = f[i1: i2 + 1], np.abs(UY[i1: i2 + 1]), np.abs(UZ[i1: i2 + 1])
    k = lambda m, n_, l: np.sqrt((m * np.pi / a) ** 2 + (n_ * np.pi / b) ** 2 + (l * np.pi / d) ** 2)
    f_TE = np.array([C0 / (2 * np.pi) * k(*m) for m in ((1, 0, 1), (1, 0, 2), (2, 0, 1), (2, 0, 2))])
    f_TM = np.array([C0 / (2 * np.pi) * k(*m) for m in ((1, 1, 0), (1, 1, 1))])
    outer = 0.02

    def outer_windows(fm):
        temp = np.concatenate(([f_start], fm, [f_stop]))
        return temp[1:] * (1 - outer), temp[:-1] * (1 + outer)
    assert check_frequency(f, UY, f_TE * (1 + 1.3e-3), f_TE * (1 - 1.3e-3), 0.6, "inside")
    assert check_frequency(f, UZ, f_TM * (1 + 0), f_TM * (1 - 2.5e-3), 0.27, "inside")
    up, lo = outer_windows(f_TE)
    assert check_frequency(f, UY, up, lo, 0.17, "outside")
    up, lo = outer_windows(f_TM)
    assert check_frequency(f, UZ, up, lo, 0.17, "outside")


def test_uniform_mesh_closed_forms():
    """dT = Delta/(c0 sqrt3) from operator.cpp:1983-2009 with float EC values,
    vi = dT/(eps0 Delta), iv = dT/(mue0 Delta)"""
    delta = 1e-3
    s = cases.uniform_box(n=(12, 13, 14), delta=delta, bc=(BC_PEC,) * 6)
    C = float(np.float32(EPS0 * delta))
    L = float(np.float32(MUE0 * delta))
    dT = 2 / np.sqrt(12 / (L * C))
    assert s.dT == pytest.approx(dT, rel=1e-14)
    assert s.dT == pytest.approx(delta / (C0 * np.sqrt(3)), rel=1e-6)
    vi = s.coeff("vi")
    iv = s.coeff("iv")
    assert vi[0, 5, 5, 5] == np.float32(dT / C)
    assert iv[1, 5, 5, 5] == np.float32(dT / L)
    # PEC faces: tangential vv/vi zero on lower faces, all comps on upper faces (operator.cpp:1113-1134)
    vv = s.coeff("vv")
    assert vv[1, 0, 5, 5] == 0 and vv[2, 0, 5, 5] == 0 and vv[0, 0, 5, 5] == 1
    assert np.all(vv[:, -1, :, :] == 0)
    # last current line always zero (operator.cpp:1176-1183)
    ii = s.coeff("ii")
    assert np.all(ii[:, -1] == 0) and np.all(ii[:, :, -1] == 0) and np.all(ii[:, :, :, -1] == 0)


@pytest.mark.parametrize("threads", [1, 3])
def test_cross_engine_bit_equality(threads):
    """TESTSUITE/enginetests/cavity.m:155: every E/H value equal between engine variants.
    scalar restatement (engine.cpp) vs sse-compressed multithreaded restatement."""
    s1 = cases.engine_cavity()
    s2 = cases.engine_cavity()
    sse = OracleSSE(s2, threads=threads)
    assert sse.unique > 1
    for n in (1, 7, 200, 292):
        s1.iterate(n)
        sse.iterate(n)
        v2, c2 = sse.fields()
        assert np.array_equal(s1.volt.view(np.uint32), v2.view(np.uint32))
        assert np.array_equal(s1.curr.view(np.uint32), c2.view(np.uint32))
    assert np.abs(s1.volt).max() > 0
    assert s1.num_ts == sse.num_ts == 500
    sse.close()


def test_cross_engine_bit_equality_odd_sizes_all_pml():
    s1 = cases.uniform_box(n=(23, 26, 21), bc=(BC_PML,) * 6, pml=4)
    s2 = cases.uniform_box(n=(23, 26, 21), bc=(BC_PML,) * 6, pml=4)
    sse = OracleSSE(s2, threads=4)
    s1.iterate(150)
    sse.iterate(150)
    v2, c2 = sse.fields()
    assert np.array_equal(s1.volt.view(np.uint32), v2.view(np.uint32))
    assert np.array_equal(s1.curr.view(np.uint32), c2.view(np.uint32))
    assert np.abs(s1.curr).max() > 0


def test_probe_equals_dump():
    """TESTSUITE/probes/fieldprobes.m:33-34: field probe == dump value at the same node (1e-7)"""
    s = cases.uniform_box(n=(20, 18, 22), bc=(BC_MUR,) * 6)
    s.iterate(60)
    pos = (11, 9, 12)
    for is_H in (0, 1):
        probe = s.raw_field(is_H, pos)
        dump = s.dump_field(is_H, 0, (8, 6, 9), (13, 12, 15))
        got = dump[:, pos[2] - 9, pos[1] - 6, pos[0] - 8]
        assert np.abs(probe).max() > 0
        assert np.allclose(got, probe, rtol=1e-7, atol=0)


def test_pml_absorbs_radiated_energy():
    """the radiated pulse is absorbed by PML_8 but stays inside a PEC box.  The soft source
    leaves a static charge (pure E) field behind, so the magnetic energy is compared."""
    fc = C0 / (20 * 1e-3) / 2
    e = {}
    for name, bc in (("pml", (BC_PML,) * 6), ("pec", (BC_PEC,) * 6)):
        s = cases.uniform_box(n=(30, 30, 30), bc=bc, pml=8, f0=fc, fc=fc)
        s.iterate(600)
        e[name] = float((s.curr.astype(np.float64) ** 2).sum() * MUE0)
    assert e["pec"] > 0 and e["pml"] < 1e-3 * e["pec"]


def test_fd_dump_restatement_matches_numpy_complex64():
    """ProcessFieldsFD::Process (processfields_fd.cpp:84-100): the restated weight and the
    complex<float> accumulation against an independent complex64 evaluation"""
    from oracle.pyoracle import OracleSim
    rng = np.random.default_rng(3)
    td = rng.standard_normal((3, 4, 5, 6)).astype(np.float32)
    acc = np.zeros(td.shape, np.complex64)
    ref = np.zeros(td.shape, np.complex64)
    dT, interval = 1.7e-12, 4
    for it in range(1, 30):
        T = it * interval * dT
        w = OracleSim.fd_weight(3.3e9, T, dT, interval)
        e = np.exp(np.complex64(-2j * np.pi * 3.3e9 * T))
        assert abs(w - e * np.float32(2) * np.float32(dT * interval)) <= 2e-7 * abs(w)
        OracleSim.fd_accumulate(acc, td, w)
        ref = (ref + (td * np.float32(w.real) + 1j * (td * np.float32(w.imag))).astype(np.complex64)).astype(np.complex64)
    assert np.array_equal(acc.view(np.uint32), ref.view(np.uint32))


def test_absorbing_sheet_restatement_absorbs():
    """Engine_Ext_Absorbing_BC restated (no reference fixture exists for it): physical pin -- the sheets of
    python/Tests/Rect_Waveguide_W_Local_Absorbers.py in small remove the pulse from a TEM line that a
    PEC-terminated line keeps, with and without super-absorption"""
    from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, EXC_E_SOFT
    x, y, z = np.arange(8) * 1e-3, np.arange(8) * 1e-3, np.arange(80) * 1e-3

    def run(sheets):
        s = OracleSim(x, y, z, 1.0)
        s.set_bc([BC_PEC, BC_PEC, BC_PMC, BC_PMC, BC_PEC, BC_PEC])
        s.set_excite_gauss(8e9, 3e9)
        s.add_excitation((x[0], y[0], z[40]), (x[-1], y[-1], z[40]), EXC_E_SOFT, (1, 0, 0))
        for a in sheets:
            s.add_absorbing_sheet(*a)
        s.build()
        s.iterate(250)
        e0 = s.energy()
        s.iterate(350)
        return e0, s.energy()
    e0, e1 = run([])
    assert e1 > 0.5 * e0
    for ty in (1, 2):
        a0, a1 = run([((0, 0, 2), (7, 7, 2), True, ty), ((0, 0, 77), (7, 7, 77), False, ty)])
        assert a0 > 0.5 * e0 and a1 < 1e-4 * a0
        k = None


def test_tfsf_restatement_injects_a_plane_wave():
    """Operator_Ext_TFSF / Engine_Ext_TFSF restated (frequency <= 0 tables): physical pin -- unit E amplitude
    inside the box (edge voltage = E * 1 mm), less than 1 % of it scattered outside, axial incidence"""
    from oracle.pyoracle import OracleSim, BC_PML
    x = y = z = np.arange(36) * 1e-3
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_PML] * 6, (6,) * 6)
    s.set_excite_gauss(0.0, 6e9)
    s.set_tfsf((9, 9, 9), (26, 26, 26), (0, 0, 1), (1, 0, 0))
    s.build()
    peak_in, peak_out = 0.0, 0.0
    for _ in range(30):
        s.iterate(8)
        v = s.volt
        peak_in = max(peak_in, float(np.abs(v[0, 12:24, 12:24, 12:24]).max()))
        peak_out = max(peak_out, float(np.abs(v[:, 7:9]).max()), float(np.abs(v[:, :, :, 28:30]).max()))
    assert 0.9e-3 < peak_in < 1.1e-3 and peak_out < 0.01 * peak_in


from tests.golden.make_golden import CASES as GOLDEN_CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_reproduces_committed_golden_vectors(name):
    """tests/golden/*.npz were written by tests/golden/make_golden.py FROM THE REFERENCE (oracle/_ref: the
    unmodified reference TUs, multithreaded engine): timestep, voltage series of every timestep and E/H bit
    digests.  The restatement must reproduce them bit for bit; the GPU suite checks the CUDA engine against the
    same files."""
    import os
    from tests.golden import make_golden as G
    g = np.load(os.path.join(os.path.dirname(G.__file__), name + ".npz"))
    assert b"libopenems_ref" in bytes(g["source"])
    s = GOLDEN_CASES[name][0]()
    assert s.dT == float(g["dT"])
    probes = [tuple(map(tuple, p)) for p in g["probes"].tolist()]
    series, digests = G.record(s, int(g["steps"]), probes)
    assert np.array_equal(series.view(np.uint64), g["series"].view(np.uint64))
    assert np.array_equal(digests, g["digests"])
