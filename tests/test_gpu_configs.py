"""BASELINE.json configs C1-C4 on the GPU engine against the oracle: probe series rel-L2
(bar 1e-5), S-parameters (bar 0.01 dB), NF2FF dumps (element-wise equal)."""
import numpy as np
import pytest

from tests import configs
from tests.gpu_util import operator_from_oracle, assert_fields_equal

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    n = np.sqrt((b ** 2).sum())
    return np.sqrt(((a - b) ** 2).sum()) / n if n else np.sqrt(((a - b) ** 2).sum())


def add_probes(eng, spec):
    for a, b in spec.get("volt", []):
        eng.AddVoltageProbe(a, b)
    for (a, b), nd, si, ei in spec.get("curr", []):
        eng.AddCurrentProbe(a, b, nd, si, ei)
    for h, p in spec.get("field", []):
        eng.AddFieldProbe(h, p)


def oracle_row(s, spec):
    row = [s.voltage_integral(a, b) for a, b in spec.get("volt", [])]
    row += [s.current_integral(a, b, nd, si, ei) for (a, b), nd, si, ei in spec.get("curr", [])]
    for h, p in spec.get("field", []):
        row += [float(v) for v in (s.curr if h else s.volt)[:, p[0], p[1], p[2]]]
    return row


@pytest.mark.parametrize("excite", ["gauss", "sinus"])
def test_c1_parallel_plate_waveguide(excite):
    s, probes = configs.c1_parallel_plate_waveguide(excite)
    eng = operator_from_oracle(s).CreateEngine()
    add_probes(eng, probes)
    interval = max(1, min(s.nyquist, 10 ** 6) // 4)
    eng.RecordProbes(interval, 400)
    ref = []
    for _ in range(40):
        s.iterate(interval)
        eng.IterateTS(interval)
        ref.append(oracle_row(s, probes))
    ts, series = eng.ReadProbeSeries()
    ref = np.array(ref)
    assert np.abs(ref[:, 0]).max() > 0
    for c in range(ref.shape[1]):
        assert rel_l2(series[:, c], ref[:, c]) <= 1e-5
    assert_fields_equal(eng, s, "C1 " + excite)


def test_c2_msl_notch_filter_sparams():
    s, ports = configs.c2_msl_notch_filter()
    eng = operator_from_oracle(s).CreateEngine()
    for p in ports:
        add_probes(eng, p)
    interval = max(1, s.nyquist // 4)
    nsamp = 300
    eng.RecordProbes(interval, nsamp)
    ref = []
    for _ in range(nsamp):
        s.iterate(interval)
        row = []
        for p in ports:
            row += oracle_row(s, p)
        ref.append(row)
    eng.IterateTS(interval * nsamp)
    ts, series = eng.ReadProbeSeries()
    ref = np.array(ref)
    assert series.shape == ref.shape and np.abs(ref).max() > 0
    for c in range(ref.shape[1]):
        assert rel_l2(series[:, c], ref[:, c]) <= 1e-5
    # S-parameters from both series (ports.py math): bar 0.01 dB
    t_u = ts * s.dT
    t_i = (ts + 0.5) * s.dT
    freq = np.linspace(1e9, 6e9, 26)

    def sparams(data):
        out = []
        for k, p in enumerate(ports):
            u = [data[:, 5 * k + q] for q in range(3)]
            i = [data[:, 5 * k + 3 + q] for q in range(2)]
            out.append(configs.msl_port_spectra(p, u, i, t_u, t_i, freq))
        s11 = out[0][1] / out[0][0]
        s21 = out[1][1] / out[0][0]
        return 20 * np.log10(np.abs(s11)), 20 * np.log10(np.abs(s21))
    g11, g21 = sparams(series)
    r11, r21 = sparams(ref)
    assert np.all(np.isfinite(r11)) and np.all(np.isfinite(r21))
    assert np.abs(g11 - r11).max() <= 0.01 and np.abs(g21 - r21).max() <= 0.01
    assert r21.min() < r21.max() - 3  # the stub does produce a notch-like variation


def test_c3_patch_antenna_nf2ff_dumps():
    s, port, faces = configs.c3_patch_antenna()
    eng = operator_from_oracle(s).CreateEngine()
    add_probes(eng, port)
    el = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], False) for p in range(s.N[n])] for n in range(3)]
    dl = [[s.edge_length(n, [p if a == n else 0 for a in range(3)], True) for p in range(s.N[n])] for n in range(3)]
    dumps = []
    for start, stop in faces:
        rng = [np.arange(start[a], stop[a] + 1) for a in range(3)]
        dumps.append((eng.AddDump(0, 2, rng[0], rng[1], rng[2], el, dl), eng.AddDump(1, 2, rng[0], rng[1], rng[2], el, dl)))
    assert len(dumps) == 6
    for nsteps in (150, 150):
        s.iterate(nsteps)
        eng.IterateTS(nsteps)
        assert np.array_equal(eng.ReadProbes(), np.array(oracle_row(s, port)))
        for (start, stop), (de, dh) in zip(faces, dumps):
            for is_H, d in ((0, de), (1, dh)):
                ref = s.dump_field(is_H, 2, start, stop)
                got = eng.ReadDump(d)
                assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.abs(s.dump_field(0, 2, *faces[5])).max() > 0
    assert_fields_equal(eng, s, "C3")


def test_c4_drude_block():
    s = configs.c4_drude_block()
    assert s.lorentz()[0]["count"] >= 24 ** 3
    eng = operator_from_oracle(s).CreateEngine()
    for nsteps in (1, 100, 300):
        s.iterate(nsteps)
        eng.IterateTS(nsteps)
        mv, mc = assert_fields_equal(eng, s, "C4")
    assert mv > 0
