"""GPU parity: probes, energy, dumps, Lorentz/Drude, lumped RLC, the compressed-operator path
and the slow per-cell accessors, all through the C ABI against the oracle."""
import numpy as np
import pytest

from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT
from tests import cases
from tests.gpu_util import operator_from_oracle, assert_fields_equal
from openems_b200 import SyntheticOperator, Engine_Interface_CUDA, EngineError, Operator_CUDA

pytestmark = pytest.mark.gpu
C0 = 299792458.0


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.sqrt(((a - b) ** 2).sum())
    n = np.sqrt((b ** 2).sum())
    return d / n if n else d


def test_probe_series_voltage_current_field_energy():
    """port voltage/current and probe series: north_star bar 1e-5 rel-L2, achieved: identical"""
    s = cases.engine_cavity()
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    vp = [((5, 4, 5), (9, 4, 5)), ((5, 8, 5), (5, 3, 5)), ((12, 4, 5), (12, 4, 17))]
    cp = [(((4, 3, 10), (12, 8, 10)), 2, (1, 1, 1), (1, 1, 1)), (((8, 2, 6), (8, 8, 20)), 0, (1, 1, 1), (1, 0, 1)),
          (((3, 5, 4), (20, 5, 25)), 1, (1, 1, 0), (1, 1, 1))]
    fp = [(0, (16, 6, 20)), (1, (16, 6, 20)), (0, (0, 0, 0)), (1, (25, 9, 31))]
    for a, b in vp:
        eng.AddVoltageProbe(a, b)
    for (a, b), nd, si, ei in cp:
        eng.AddCurrentProbe(a, b, nd, si, ei)
    for h, p in fp:
        eng.AddFieldProbe(h, p)
    interval = max(1, s.nyquist // 4)
    eng.RecordProbes(interval, 200)
    ref = []
    for it in range(60):
        s.iterate(interval)
        eng.IterateTS(interval)
        row = [s.voltage_integral(a, b) for a, b in vp]
        row += [s.current_integral(a, b, nd, si, ei) for (a, b), nd, si, ei in cp]
        for h, p in fp:
            v = (s.curr if h else s.volt)[:, p[0], p[1], p[2]]
            row += [float(x) for x in v]
        ref.append(row)
        if it % 20 == 7:
            now = eng.ReadProbes()
            assert np.array_equal(now, np.array(row))
            e_gpu, e_ref = eng.CalcFastEnergy(), s.energy()
            assert e_ref > 0 and abs(e_gpu - e_ref) <= 1e-12 * e_ref
    ts, series = eng.ReadProbeSeries()
    ref = np.array(ref)
    assert series.shape == ref.shape
    assert list(ts) == [interval * (i + 1) for i in range(60)]
    assert np.abs(ref).max(axis=0).min() >= 0
    for col in range(ref.shape[1]):
        assert rel_l2(series[:, col], ref[:, col]) <= 1e-5
    assert np.array_equal(series, ref)  # in fact identical
    assert np.abs(ref[:, :6]).max() > 0


def test_engine_interface_mirror():
    s = cases.uniform_box(n=(20, 18, 22), bc=(BC_MUR,) * 6)
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    ei = Engine_Interface_CUDA(op, eng)
    s.iterate(60)
    eng.IterateTS(60)
    assert ei.GetNumberOfTimesteps() == 60
    assert ei.GetTime() == pytest.approx(60 * s.dT)
    assert ei.CalcVoltageIntegral((8, 9, 11), (12, 9, 11)) == s.voltage_integral((8, 9, 11), (12, 9, 11))
    pos = (11, 9, 12)
    assert np.array_equal(ei.GetEField(pos), s.raw_field(0, pos))
    assert np.array_equal(ei.GetHField(pos), s.raw_field(1, pos))
    # per-cell slow path, Engine::GetVolt/SetVolt
    assert eng.GetVolt(2, 10, 9, 11) == s.volt[2, 10, 9, 11]
    assert eng.GetCurr(1, (10, 9, 11)) == s.curr[1, 10, 9, 11]
    eng.SetVolt(0, 3, 4, 5, 1.25)
    assert eng.GetVolt(0, (3, 4, 5)) == 1.25


@pytest.mark.parametrize("interp", [0, 1, 2])
@pytest.mark.parametrize("is_H", [0, 1])
def test_field_dump(is_H, interp):
    """probe==dump rule and bit-equal dumps (fieldprobes.m:34, enginetests/cavity.m:155) on a
    non-uniform mesh, including the faces where the interpolation degenerates"""
    x = np.cumsum(np.r_[0, np.linspace(1, 2, 17)]) * 1e-3
    y = np.arange(16) * 1.5e-3
    z = np.cumsum(np.r_[0, np.linspace(2, 1, 19)]) * 1e-3
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_MUR, BC_MUR, BC_PEC, BC_PMC, BC_MUR, BC_MUR])
    s.set_excite_gauss(5e9, 5e9)
    c = cases.edge_center((x, y, z), 2, (9, 8, 10))
    s.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
    s.build()
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    s.iterate(70)
    eng.IterateTS(70)
    start, stop = (0, 0, 0), (17, 15, 19)
    el = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], False) for p in range(s.N[n])] for n in range(3)]
    dl = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], True) for p in range(s.N[n])] for n in range(3)]
    d = eng.AddDump(is_H, interp, np.arange(18), np.arange(16), np.arange(20), el, dl)
    got = eng.ReadDump(d)
    ref = s.dump_field(is_H, interp, start, stop)
    assert np.abs(ref).max() > 0
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # sub-sampled position lists
    px, py, pz = [2, 5, 9, 17], [0, 8, 15], [3, 10, 19]
    d2 = eng.AddDump(is_H, interp, px, py, pz, el, dl)
    sub = eng.ReadDump(d2)
    assert np.array_equal(sub, ref[:, pz][:, :, py][:, :, :, px])


def test_async_dump_does_not_stall_and_matches():
    """oems_cuda_read_dump_async / oems_cuda_wait: the dump captures the fields of the timestep it was issued at; the
    engine steps on while the copy is in flight, a second dump of the same box queues behind the first"""
    s = cases.uniform_box(n=(30, 26, 28), bc=(BC_PML, BC_PML, BC_MUR, BC_MUR, BC_PEC, BC_PML), pml=5)
    eng = operator_from_oracle(s).CreateEngine()
    start, stop = (3, 2, 1), (26, 23, 25)
    el = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], False) for p in range(s.N[n])] for n in range(3)]
    dl = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], True) for p in range(s.N[n])] for n in range(3)]
    d = eng.AddDump(0, 2, np.arange(start[0], stop[0] + 1), np.arange(start[1], stop[1] + 1), np.arange(start[2], stop[2] + 1), el, dl)
    s.iterate(40)
    eng.IterateTS(40)
    ref40 = s.dump_field(0, 2, start, stop)
    t = eng.ReadDumpAsync(d)
    eng.IterateTS(25)             # enqueued behind the dump kernel, overlapping its copy
    got40 = eng.WaitDump(t)
    assert np.abs(ref40).max() > 0 and np.array_equal(got40.view(np.uint32), ref40.view(np.uint32))
    s.iterate(25)
    t1 = eng.ReadDumpAsync(d)
    eng.IterateTS(10)
    s.iterate(10)
    ref65 = s.dump_field(0, 2, start, stop)  # oracle is now at 75; recompute the 65 reference from a fresh run below
    got65 = eng.WaitDump(t1)
    s2 = cases.uniform_box(n=(30, 26, 28), bc=(BC_PML, BC_PML, BC_MUR, BC_MUR, BC_PEC, BC_PML), pml=5)
    s2.iterate(65)
    assert np.array_equal(got65.view(np.uint32), s2.dump_field(0, 2, start, stop).view(np.uint32))
    assert not np.array_equal(got65, ref65)
    assert np.array_equal(eng.ReadDump(d).view(np.uint32), ref65.view(np.uint32))
    assert_fields_equal(eng, s, "after async dumps")


def test_lorentz_drude_block():
    """config C4 in small: Drude eps+mue block (f_p 5 GHz, tau 5 ns) and a 2-pole Lorentz block"""
    n = (34, 30, 38)
    lor = [dict(start=(0.010, 0.008, 0.012), stop=(0.022, 0.020, 0.026), eps_fp=(5e9,), eps_tau=(5e-9,),
                mue_fp=(5e9,), mue_tau=(5e-9,)),
           dict(start=(0.004, 0.004, 0.004), stop=(0.008, 0.012, 0.010), epsR=2.0, eps_fp=(3e9, 6e9), eps_tau=(2e-9, 0.0),
                eps_flor=(0.0, 9e9), prio=3)]
    fc = C0 / (20 * 1e-3) / 2
    s = cases.uniform_box(n=n, bc=(BC_PML,) * 6, pml=6, f0=fc, fc=fc, lorentz=lor, src_pos=(6, 15, 19))
    L = s.lorentz()
    assert len(L) == 2 and L[0]["count"] > 1000 and (L[0]["flags"] & 3) == 3 and (L[1]["flags"] & 4)
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    for nsteps in (1, 40, 200):
        s.iterate(nsteps)
        eng.IterateTS(nsteps)
        mv, mc = assert_fields_equal(eng, s, "lorentz")
    assert mv > 0


@pytest.mark.parametrize("case", ["c4", "two_boxes", "into_pml", "next_to_mur"])
def test_lorentz_drude_in_the_one_pass_schedule(case):
    """the one-pass kernel applies the ADE of dispersive cells itself (LOR instance of k_fused_tma; the pre hooks run as
    list kernels on the source set): C4 in small (Drude eps+mue block), a Drude block plus a 2-pole Lorentz block
    (two orders); both schedules bit-equal to the oracle, switched mid-run at odd and even timestep counts.  Dispersive
    cells inside a UPML box or on a Mur plane's lines keep the two-pass schedule (the hook order could not be kept)"""
    from tests import configs
    from tests.test_gpu_parity import run_both
    fc = C0 / (20 * 1e-3) / 2
    fusable = True
    if case == "c4":
        s = configs.c4_drude_block(block=(12, 35))   # (x = 36 .. 39 is the float4 chunk the x-high UPML box starts in)
    elif case == "two_boxes":
        lor = [dict(start=(0.010, 0.008, 0.012), stop=(0.022, 0.020, 0.026), eps_fp=(5e9,), eps_tau=(5e-9,), mue_fp=(5e9,), mue_tau=(5e-9,)),
               dict(start=(0.008, 0.007, 0.007), stop=(0.012, 0.012, 0.010), epsR=2.0, eps_fp=(3e9, 6e9), eps_tau=(2e-9, 0.0),
                    eps_flor=(0.0, 9e9), prio=3)]
        s = cases.uniform_box(n=(34, 30, 38), bc=(BC_PML,) * 6, pml=6, f0=fc, fc=fc, lorentz=lor, src_pos=(6, 15, 19))
    elif case == "into_pml":
        lor = [dict(start=(0.0, 0.008, 0.012), stop=(0.012, 0.020, 0.026), eps_fp=(5e9,), eps_tau=(5e-9,))]
        s = cases.uniform_box(n=(34, 30, 38), bc=(BC_PML,) * 6, pml=6, f0=fc, fc=fc, lorentz=lor, src_pos=(20, 15, 19))
        fusable = False
    else:
        lor = [dict(start=(0.0, 0.008, 0.012), stop=(0.012, 0.020, 0.026), eps_fp=(5e9,), eps_tau=(5e-9,), mue_fp=(4e9,), mue_tau=(3e-9,))]
        s = cases.uniform_box(n=(34, 30, 38), bc=(BC_MUR, BC_PML, BC_PEC, BC_PML, BC_PMC, BC_PML), pml=6, f0=fc, fc=fc, lorentz=lor, src_pos=(20, 15, 19))
        fusable = False
    assert s.lorentz()[0]["count"] > 500
    eng = run_both(s, steps=(1, 2, 37, 120), what="dispersive " + case)
    names = [n for n, _ in eng.TimeSchedule(0)]
    assert eng.GetOption("fused") == int(fusable)
    if fusable:
        assert "fused_EH" in names and "lorentz_pre_V" in names and "lorentz_apply_V" not in names
        for steps, fused in ((3, 0), (4, 1), (5, 0), (2, 1), (30, 1)):
            eng.SetOption("fused", fused)
            assert eng.GetOption("fused") == fused
            s.iterate(steps)
            eng.IterateTS(steps)
            assert_fields_equal(eng, s, "dispersive %s, schedule switched to %d" % (case, fused))
        eng.SetOption("tma", 0)   # no ADE instance of the register-staged kernel: two-pass
        assert eng.GetOption("fused") == 0
        s.iterate(7)
        eng.IterateTS(7)
        assert_fields_equal(eng, s, "dispersive %s, tma off" % case)
    else:
        assert "lorentz_apply_V" in names and "fused_EH" not in names


def test_lumped_rlc_raw():
    rng = np.random.default_rng(7)
    n = (24, 22, 26)

    def extra(s, lines):
        cnt = 5
        pos = np.array([[8, 9, 10, 11, 12], [10, 10, 11, 11, 12], [12, 13, 12, 13, 14]], np.uint32)
        d = np.array([0, 1, 2, 2, 1], np.int32)
        co = {k: (rng.uniform(-0.3, 0.3, cnt)).astype(np.float32) for k in ("ilv", "i2v", "vv2", "vj1", "vj2", "ib0", "b1", "b2")}
        co["vvd"] = rng.uniform(0.5, 1.0, cnt).astype(np.float32)
        s._rlc = (d, pos, co)
        s.add_rlc_raw(d, pos, co)
    s = cases.uniform_box(n=n, bc=(BC_MUR,) * 6, extra=extra, src_pos=(9, 10, 12))
    op = operator_from_oracle(s)
    op.AddLumpedRLC(*s._rlc)
    eng = op.CreateEngine()
    for nsteps in (1, 2, 3, 4, 50):
        s.iterate(nsteps)
        eng.IterateTS(nsteps)
        assert_fields_equal(eng, s, "rlc")


def test_compressed_operator_path_matches_oracle():
    """host builder -> compressed upload -> kernels, against the oracle's dense pipeline"""
    lines = (np.arange(40, dtype=np.float64), np.arange(33, dtype=np.float64), np.arange(45, dtype=np.float64))

    def setup(q):
        q.set_bc([BC_PML, BC_PML, BC_MUR, BC_PML, BC_PMC, BC_PML], (8, 8, 8, 6, 8, 7))
        q.set_excite_gauss(6e9, 6e9)
        q.add_material((10, 5, 8), (25, 20, 30), epsR=2.5, kappa=0.01)
        q.add_metal((12, 10, 20), (30, 18, 20))
        q.add_excitation((20, 16, 10.5), (20, 16, 10.5), EXC_E_SOFT, (0, 0, 1))
    o = OracleSim(*lines, 1e-3)
    p = SyntheticOperator(*lines, 1e-3)
    setup(o)
    setup(p)
    o.build()
    p.build()
    eng = p.CreateEngine()
    for nsteps in (1, 30, 170):
        o.iterate(nsteps)
        eng.IterateTS(nsteps)
        mv, mc = assert_fields_equal(eng, o, "compressed path")
    assert mv > 0 and mc > 0
    st = eng.GetStats()
    assert st["n_unique"] == p.n_unique and st["index_bytes"] == 2 and st["uses_graph"]


def test_reset_and_error_behaviour():
    s = cases.uniform_box(n=(16, 16, 16), bc=(BC_PEC,) * 6)
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    eng.IterateTS(20)
    assert np.abs(eng.GetFields(0)).max() > 0
    eng.Reset()
    assert eng.GetNumberOfTimesteps() == 0 and np.abs(eng.GetFields(0)).max() == 0
    s.iterate(25)
    eng.IterateTS(25)
    assert_fields_equal(eng, s, "after reset")
    with pytest.raises(EngineError):
        eng.GetVolt(0, 99, 0, 0)
    with pytest.raises(EngineError):
        eng.AddVoltageProbe((0, 0, 0), (0, 0, 99))
    with pytest.raises(EngineError):
        Operator_CUDA((2, 5, 5)).CreateEngine()
    bad = Operator_CUDA((8, 8, 8))
    with pytest.raises(EngineError):
        bad.CreateEngine()  # no coefficients


def test_upml_cells_outside_h_update_follow_reference():
    """Engine_Ext_UPML touches every cell of its box, also the last H line the stencil skips
    (engine_ext_upml.cpp:63-90 vs engine.cpp:179-183): poke a current there and compare"""
    s = cases.uniform_box(n=(20, 18, 22), bc=(BC_PML,) * 6, pml=4)
    op = operator_from_oracle(s)
    eng = op.CreateEngine()
    assert "upml_untouched_H" not in [x for x, _ in eng.TimeSchedule(0)]
    for n, pos in ((1, (19, 5, 6)), (0, (7, 17, 3)), (2, (4, 9, 21))):
        s.curr[n, pos[0], pos[1], pos[2]] = 0.37
        eng.SetCurr(n, pos, 0.37)
    names = [x for x, _ in eng.TimeSchedule(0)]
    assert "upml_untouched_H" in names and "fused_EH" not in names  # the edge kernel is now scheduled (two-pass)
    for nsteps in (1, 1, 10):
        s.iterate(nsteps)
        eng.IterateTS(nsteps)
        assert_fields_equal(eng, s, "poked last-line currents")


@pytest.mark.parametrize("fused", [0, 1])
def test_steady_state_detection_sinus(fused):
    """Engine_Ext_SteadyState with the stock sinus-excited parallel plate
    waveguide (C1): the device-recorded criterion equals the reference's, period by period"""
    from tests import configs
    s0, _ = configs.c1_parallel_plate_waveguide("sinus")
    sv, si, period = s0.signal()
    assert period > 0
    op = operator_from_oracle(s0)
    op.SetSteadyStateDetection(period)
    per, pos3, d = op.steadystate
    # the oracle with the same probe set (it must be added before build -> rebuild the case)
    from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, BC_MUR, EXC_E_SOFT
    s = OracleSim(s0.x, s0.y, s0.z, 1.0)
    s.set_bc([BC_PMC, BC_PMC, BC_PEC, BC_PEC, BC_MUR, BC_MUR])
    s.set_excite_sinus(10e6)
    s.add_excitation((-10, -10, 0), (10, 10, 0), EXC_E_SOFT, (0, 1, 0))
    s.add_steadystate(per, pos3, d.astype(np.int32))
    s.build()
    eng = op.CreateEngine()
    eng.SetOption("fused", fused)
    checks = 0
    for it in range(9):
        n = per if it else per + 1   # stop right after the timestep with TS % period == 0
        s.iterate(n)
        eng.IterateTS(n)
        got, checks = eng.SteadyStateLastDiff()
        ref = s.steadystate_last_diff()
        assert got == pytest.approx(ref, rel=1e-9, abs=1e-300), (it, got, ref)
    assert checks >= 7 and 0 < ref < 1
    assert_fields_equal(eng, s, "steady state run")


@pytest.mark.parametrize("is_H,interp", [(0, 2), (1, 2), (0, 0)])
def test_fd_dump_running_dft_on_device(is_H, interp):
    """ProcessFieldsFD (processfields_fd.cpp:72-107) -- the running DFT of an
    NF2FF-style dump box accumulated on the device, bit-equal to the reference's
    complex<float> accumulation of the same samples; D2H only at the end"""
    s = cases.uniform_box(n=(30, 26, 28), bc=(BC_PML, BC_PML, BC_MUR, BC_MUR, BC_PEC, BC_PML), pml=5)
    eng = operator_from_oracle(s).CreateEngine()
    start, stop = (7, 6, 3), (22, 19, 21)
    el = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], False) for p in range(s.N[n])] for n in range(3)]
    dl = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], True) for p in range(s.N[n])] for n in range(3)]
    d = eng.AddDump(is_H, interp, np.arange(start[0], stop[0] + 1), np.arange(start[1], stop[1] + 1),
                    np.arange(start[2], stop[2] + 1), el, dl)
    freqs = [2e9, 5.5e9, 7.5e9]
    fd = eng.AddFDDump(d, len(freqs))
    interval = 3
    ref = np.zeros((len(freqs),) + (3, stop[2] - start[2] + 1, stop[1] - start[1] + 1, stop[0] - start[0] + 1), np.complex64)
    for it in range(40):
        s.iterate(interval)
        eng.IterateTS(interval)
        # Engine_Interface_FDTD::GetTime(dualTime): H dumps are half a timestep later
        T = (s.num_ts + (0.5 if is_H else 0.0)) * s.dT
        w = np.array([OracleSim.fd_weight(f, T, s.dT, interval) for f in freqs], np.complex64)
        eng.AccumulateFD(fd, w)
        td = s.dump_field(is_H, interp, start, stop)
        for n in range(len(freqs)):
            OracleSim.fd_accumulate(ref[n], td, w[n])
    got, nsamp = eng.ReadFD(fd)
    assert nsamp == 40 and np.abs(ref).max() > 0
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    with pytest.raises(EngineError):
        eng.AddFDDump(d, 0)
    with pytest.raises(EngineError):
        eng.AddFDDump(99, 1)


@pytest.mark.parametrize("is_H", [0, 1])
def test_mode_matching_plane_integral(is_H):
    """ProcessModeMatch::CalcMultipleIntegrals (processmodematch.cpp:222-266) on
    the device -- TE10-like template on a cross-section of a PEC waveguide (graded mesh), value and
    purity equal to the reference's sequential fp64 sums at every sample"""
    rng = np.random.default_rng(5)
    x = np.cumsum(np.r_[0, 1 + 0.3 * rng.random(23)]) * 1e-3
    y = np.cumsum(np.r_[0, 1 + 0.2 * rng.random(15)]) * 1e-3
    z = np.arange(60) * 1e-3
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_PEC, BC_PEC, BC_PEC, BC_PEC, BC_PML, BC_PML], (6,) * 6)
    s.set_excite_gauss(9e9, 3e9)
    s.add_excitation((x[0], y[0], z[12]), (x[-1], y[-1], z[12]), EXC_E_SOFT, (0, 1, 0))
    s.build()
    eng = operator_from_oracle(s).CreateEngine()
    ny = 2
    # ProcessModeMatch::InitProcess pulls the surface off the boundaries (lines 97-100)
    start, stop = (1, 1, 40), (len(x) - 2, len(y) - 2, 40)
    el = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], False) for p in range(s.N[n])] for n in range(3)]
    dl = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], True) for p in range(s.N[n])] for n in range(3)]
    nl0, nl1 = stop[0] - start[0] + 1, stop[1] - start[1] + 1
    dist = np.zeros((2, nl0, nl1))
    area = np.zeros((nl0, nl1))
    for a in range(nl0):
        for b in range(nl1):
            pos = [start[0] + a, start[1] + b, start[2]]
            xx = s.disc_line(0, pos[0], bool(is_H))
            # E mode: Ey ~ sin(pi x / a); H mode: Hx ~ -sin(pi x / a)
            dist[0, a, b] = -np.sin(np.pi * xx / x[-1]) if is_H else 0.0
            dist[1, a, b] = 0.0 if is_H else np.sin(np.pi * xx / x[-1])
            area[a, b] = s.edge_length(0, pos, not is_H) * s.edge_length(1, pos, not is_H)
    dist /= np.sqrt(((dist ** 2) * area).sum())
    m = eng.AddModeMatch(is_H, ny, start, stop, dist[0], dist[1], area, el, dl)
    peak = 0.0
    for it in range(60):
        s.iterate(5)
        eng.IterateTS(5)
        got = eng.ReadModeMatch(m)
        ref = s.mode_match(is_H, ny, start, stop, dist[0], dist[1])
        assert got == ref, (it, got, ref)
        peak = max(peak, abs(ref[0]))
    assert peak > 0 and 0.5 < ref[1] <= 1.0 + 1e-12  # the excited field is mostly the template mode
