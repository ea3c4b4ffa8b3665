"""Readout and the rank-(f) extensions on z-slab engines (SURVEY 8e "extension sharding"): field dumps with all
interpolation types, FD dumps, mode matching, local absorbing sheets and steady-state detection on 3 slabs equal
the single-domain oracle.  The slabs live on ONE GPU here (the multi-process IPC path runs the same engine code,
tests/test_gpu_ipc.py); every slab evaluates the planes it owns after the neighbours have completed its ghost
planes (oems_cuda_exchange_ghosts), the host concatenates / adds the pieces (openems_b200/slabs.py)."""
import numpy as np
import pytest

from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT
from tests import cases
from tests.gpu_util import operator_from_oracle
from openems_b200 import slabs
from openems_b200.slabs import held_range, link_engines_in_process

pytestmark = pytest.mark.gpu


def make_slabs(s, bounds, fused=None, op=None):
    op = op or operator_from_oracle(s)
    engines = [op.CreateEngine(device=0, slab=(bounds[r], bounds[r + 1])) for r in range(len(bounds) - 1)]
    if fused is not None:
        for e in engines:
            e.SetOption("fused", fused)
    link_engines_in_process(engines)
    return engines


def step(engines, s, n):
    s.iterate(n)
    for _ in range(n):
        for e in engines:
            e.IterateTS(1)


def fields_equal(engines, s, bounds, what):
    nz = s.N[2]
    for e in engines:
        e.Synchronize()
    for w, ref in ((0, s.volt), (1, s.curr)):
        got = np.zeros_like(ref)
        for r, e in enumerate(engines):
            zb, ze = bounds[r], bounds[r + 1]
            h0, _ = held_range(nz, zb, ze)
            got[..., zb:ze] = e.GetFields(w)[..., zb - h0: ze - h0]
        bad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
        assert bad == 0, "%s: %d values of field %d differ" % (what, bad, w)


def edge_tables(s):
    el = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], False) for p in range(s.N[n])] for n in range(3)]
    dl = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], True) for p in range(s.N[n])] for n in range(3)]
    return el, dl


@pytest.mark.parametrize("fused", [0, 1])
def test_field_dumps_on_slabs(fused):
    """all six dump variants (E/H x no/node/cell interpolation) over the whole non-uniform mesh, including the slab
    interfaces where the interpolation reads the neighbour's plane, twice (the second exchange follows a release),
    plus a sub-sampled box that leaves one slab without any line"""
    x = np.cumsum(np.r_[0, np.linspace(1, 2, 17)]) * 1e-3
    y = np.arange(16) * 1.5e-3
    z = np.cumsum(np.r_[0, np.linspace(2, 1, 19)]) * 1e-3
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_MUR, BC_MUR, BC_PEC, BC_PMC, BC_MUR, BC_MUR])
    s.set_excite_gauss(5e9, 5e9)
    c = cases.edge_center((x, y, z), 2, (9, 8, 10))
    s.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
    s.build()
    bounds = [0, 7, 13, 20]
    engines = make_slabs(s, bounds, fused)
    assert all(e.GetOption("fused") == fused for e in engines)
    el, dl = edge_tables(s)
    ids = {}
    for is_H in (0, 1):
        for interp in (0, 1, 2):
            ids[(is_H, interp)] = [e.AddDump(is_H, interp, np.arange(18), np.arange(16), np.arange(20), el, dl) for e in engines]
    pz = [3, 6, 14, 15, 19]   # nothing between 7 and 12: the middle slab owns no line of this box
    sub = [e.AddDump(0, 1, [2, 5, 17], [0, 8, 15], pz, el, dl) for e in engines]
    assert [e.DumpOwnRange(d) for e, d in zip(engines, sub)] == [(0, 2), (2, 0), (2, 3)]
    for n in (70, 1, 24):
        step(engines, s, n)
        for (is_H, interp), d in ids.items():
            got = slabs.read_dump_slabs(engines, d)
            ref = s.dump_field(is_H, interp, (0, 0, 0), (17, 15, 19))
            assert np.abs(ref).max() > 0
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (is_H, interp, s.num_ts)
        got = slabs.read_dump_slabs(engines, sub)
        ref = s.dump_field(0, 1, (0, 0, 0), (17, 15, 19))[:, pz][:, :, [0, 8, 15]][:, :, :, [2, 5, 17]]
        assert np.array_equal(got, ref)
    fields_equal(engines, s, bounds, "after dumps")   # the ghost exchange leaves the time loop's values alone


def test_fd_dump_on_slabs():
    """running DFT of an NF2FF-style box accumulated per slab on the device (cell interpolation reads the plane above)"""
    s = cases.uniform_box(n=(30, 26, 28), bc=(BC_PML, BC_PML, BC_MUR, BC_MUR, BC_PEC, BC_PML), pml=5)
    bounds = [0, 10, 19, 28]
    engines = make_slabs(s, bounds)
    start, stop = (7, 6, 3), (22, 19, 21)
    el, dl = edge_tables(s)
    freqs = [2e9, 5.5e9, 7.5e9]
    interval = 3
    for is_H, interp in ((0, 2), (1, 1)):
        d = [e.AddDump(is_H, interp, np.arange(start[0], stop[0] + 1), np.arange(start[1], stop[1] + 1),
                       np.arange(start[2], stop[2] + 1), el, dl) for e in engines]
        fd = [e.AddFDDump(i, len(freqs)) for e, i in zip(engines, d)]
        ref = np.zeros((len(freqs), 3, stop[2] - start[2] + 1, stop[1] - start[1] + 1, stop[0] - start[0] + 1), np.complex64)
        for it in range(25):
            step(engines, s, interval)
            T = (s.num_ts + (0.5 if is_H else 0.0)) * s.dT
            w = np.array([OracleSim.fd_weight(f, T, s.dT, interval) for f in freqs], np.complex64)
            slabs.accumulate_fd_slabs(engines, fd, w)
            td = s.dump_field(is_H, interp, start, stop)
            for n in range(len(freqs)):
                OracleSim.fd_accumulate(ref[n], td, w[n])
        got, nsamp = slabs.read_fd_slabs(engines, fd)
        assert nsamp == 25 and np.abs(ref).max() > 0
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("is_H", [0, 1])
def test_mode_matching_on_slabs(is_H):
    """a cross-section normal to z that is the LOWEST owned plane of a slab (node interpolation reads the neighbour's
    plane): equal to the oracle; a plane normal to x that crosses all slabs: partial sums added on the host (1e-12)"""
    rng = np.random.default_rng(5)
    x = np.cumsum(np.r_[0, 1 + 0.3 * rng.random(23)]) * 1e-3
    y = np.cumsum(np.r_[0, 1 + 0.2 * rng.random(15)]) * 1e-3
    z = np.arange(60) * 1e-3
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_PEC, BC_PEC, BC_PEC, BC_PEC, BC_PML, BC_PML], (6,) * 6)
    s.set_excite_gauss(9e9, 3e9)
    s.add_excitation((x[0], y[0], z[12]), (x[-1], y[-1], z[12]), EXC_E_SOFT, (0, 1, 0))
    s.build()
    bounds = [0, 20, 40, 60]
    engines = make_slabs(s, bounds)
    el, dl = edge_tables(s)

    def template(ny, start, stop):
        nP, nPP = (ny + 1) % 3, (ny + 2) % 3
        nl0, nl1 = stop[nP] - start[nP] + 1, stop[nPP] - start[nPP] + 1
        dist = np.zeros((2, nl0, nl1))
        area = np.zeros((nl0, nl1))
        for a in range(nl0):
            for b in range(nl1):
                pos = [0, 0, 0]
                pos[ny] = start[ny]; pos[nP] = start[nP] + a; pos[nPP] = start[nPP] + b
                xx = s.disc_line(0, pos[0], bool(is_H)) if ny == 2 else s.disc_line(1, pos[1], bool(is_H))
                L = x[-1] if ny == 2 else y[-1]
                dist[0, a, b] = np.sin(np.pi * xx / L) * (0.3 if ny == 2 else 1.0)
                dist[1, a, b] = np.cos(np.pi * xx / L)
                area[a, b] = s.edge_length(nP, pos, not is_H) * s.edge_length(nPP, pos, not is_H)
        dist /= np.sqrt(((dist ** 2) * area).sum())
        return dist, area

    planes = []
    for ny, start, stop in ((2, (1, 1, 40), (len(x) - 2, len(y) - 2, 40)), (0, (10, 1, 8), (10, len(y) - 2, 52))):
        dist, area = template(ny, start, stop)
        planes.append((ny, start, stop, dist, [e.AddModeMatch(is_H, ny, start, stop, dist[0], dist[1], area, el, dl) for e in engines]))
    peak = 0.0
    for it in range(30):
        step(engines, s, 7)
        for ny, start, stop, dist, m in planes:
            got = slabs.mode_match_slabs(engines, m)
            ref = s.mode_match(is_H, ny, start, stop, dist[0], dist[1])
            if ny == 2:
                assert got == ref, (it, got, ref)
            else:
                assert got[0] == pytest.approx(ref[0], rel=1e-12, abs=1e-30) and got[1] == pytest.approx(ref[1], rel=1e-11, abs=1e-30), (it, got, ref)
            peak = max(peak, abs(ref[0]))
    assert peak > 0


@pytest.mark.parametrize("fused", [0, 1])
def test_absorbing_sheets_on_slabs(fused):
    """local absorbing sheets (Engine_Ext_Absorbing_BC) on 3 slabs: a y-normal and an x-normal sheet cut by slab
    interfaces, a super-absorbing z-normal sheet inside one slab; fields bit-equal to the single-domain oracle"""
    x, y, z = np.arange(34) * 1e-3, np.cumsum(np.r_[0, np.linspace(1, 1.6, 27)]) * 1e-3, np.arange(30) * 1e-3
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_PML, BC_MUR, BC_PEC, BC_PML, BC_PMC, BC_PEC], (5,) * 6)
    s.set_excite_gauss(6e9, 4e9)
    c = cases.edge_center((x, y, z), 2, (15, 12, 14))
    s.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
    s.add_absorbing_sheet((8, 2, 20), (25, 2, 27), True, 1, 0.0)        # y-normal, plain Mur, cut at z = 22
    s.add_absorbing_sheet((3, 3, 27), (30, 20, 27), False, 2, 0.0)      # z-normal, super-absorbing: lines 25..27 in the top slab
    s.add_absorbing_sheet((29, 4, 3), (29, 18, 12), False, 2, 2.5e8)    # x-normal, cut at z = 10
    s.build()
    bounds = [0, 10, 22, 30]
    engines = make_slabs(s, bounds, fused)
    assert all(e.GetOption("fused") == fused for e in engines)
    for n in (1, 2, 25, 90):
        step(engines, s, n)
        fields_equal(engines, s, bounds, "absorbing sheets on slabs, %d steps" % s.num_ts)
    assert np.abs(s.volt).max() > 0
    # a z-normal sheet on a slab interface is refused with a clear message
    from openems_b200 import EngineError
    with pytest.raises(EngineError, match="slab boundary"):
        make_slabs(s, [0, 10, 26, 30], fused)


def test_steady_state_on_slabs():
    """Engine_Ext_SteadyState on 3 slabs: each slab records its probes and the energy of its planes, the host puts
    them together; the criterion equals the oracle's period by period"""
    from tests import configs
    s0, _ = configs.c1_parallel_plate_waveguide("sinus")
    sv, si, period = s0.signal()
    op = operator_from_oracle(s0)
    op.SetSteadyStateDetection(period)
    per, pos3, d = op.steadystate
    s = OracleSim(s0.x, s0.y, s0.z, 1.0)
    s.set_bc([BC_PMC, BC_PMC, BC_PEC, BC_PEC, BC_MUR, BC_MUR])
    s.set_excite_sinus(10e6)
    s.add_excitation((-10, -10, 0), (10, 10, 0), EXC_E_SOFT, (0, 1, 0))
    s.add_steadystate(per, pos3, d.astype(np.int32))
    s.build()
    nz = s.N[2]
    bounds = [0, nz // 3, 2 * nz // 3 + 1, nz]
    engines = make_slabs(s, bounds, op=op)
    # (the criterion is a maximum over the probes: the order in which the slabs' records are concatenated is free)
    checks = 0
    for it in range(7):
        step(engines, s, per if it else per + 1)
        got, checks = slabs.steadystate_slabs(engines, per)
        ref = s.steadystate_last_diff()
        assert got == pytest.approx(ref, rel=1e-9, abs=1e-300), (it, got, ref)
    assert checks >= 5 and 0 < ref < 1
    fields_equal(engines, s, bounds, "steady state on slabs")


def test_c3_patch_antenna_nf2ff_on_two_slabs():
    """BASELINE config C3 on 2 z-slabs: the 12 NF2FF dumps (6 faces x E/H, cell interpolation) and the port probes
    equal the single-domain oracle element-wise; PML_8 on all faces, lumped port, the patch metal"""
    from tests import configs
    from tests.test_gpu_configs import add_probes, oracle_row
    s, port, faces = configs.c3_patch_antenna()
    nz = s.N[2]
    bounds = [0, nz // 2 + 1, nz]
    engines = make_slabs(s, bounds)
    for e in engines:
        add_probes(e, port)
    el, dl = edge_tables(s)
    dumps = []
    for start, stop in faces:
        rng = [np.arange(start[a], stop[a] + 1) for a in range(3)]
        dumps.append(([e.AddDump(0, 2, rng[0], rng[1], rng[2], el, dl) for e in engines],
                      [e.AddDump(1, 2, rng[0], rng[1], rng[2], el, dl) for e in engines]))
    for nsteps in (150, 150):
        step(engines, s, nsteps)
        # probes spanning slabs: every slab sums its terms, the host adds the partial sums
        got = np.sum([e.ReadProbes() for e in engines], axis=0)
        ref = np.array(oracle_row(s, port))
        assert np.allclose(got, ref, rtol=1e-12, atol=0)
        for (start, stop), (de, dh) in zip(faces, dumps):
            for is_H, d in ((0, de), (1, dh)):
                ref = s.dump_field(is_H, 2, start, stop)
                got = slabs.read_dump_slabs(engines, d)
                assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (is_H, start, stop)
    assert np.abs(s.dump_field(0, 2, *faces[5])).max() > 0
    fields_equal(engines, s, bounds, "C3 on two slabs")
