"""The north_star end to end: the REFERENCE'S OWN operator build (Operator_Multithread::CalcECOperator and every
operator extension), its own Processing classes and its own RunFDTD loop drive the B200 engine through
integration/*.cpp (Operator_CUDA, Engine_CUDA, Engine_Interface_CUDA_FDTD, Process*_CUDA) and the C ABI -- and
every probe file, dump and field equals what the reference's multithreaded sse-compressed CPU engine produces
from the same setup.  Both run inside oracle/_ref/libopenems_ref_cuda.so (the unmodified reference translation
units + integration/ + libopenems_b200.so).  Correctness bar of BASELINE.json: 1e-5 rel-L2; measured: 0 (identical
files, identical bits)."""
import os

import numpy as np
import pytest

from oracle import pyref
from oracle.pyoracle import BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT
from oracle.pyref import RefSim, ENGINE_MULTITHREADED, ENGINE_CUDA, ENGINE_BASIC
from tests import cases, configs
from tests.ref_util import backend, ref_class, assert_same, assert_operator_equal

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(pyref._LIB_CUDA) and not pyref.have_reference_tree(),
                                                  reason="oracle/_ref/libopenems_ref_cuda.so not available")]
C0 = 299792458.0


def pair(fn, *a, fast=True, cpu_engine=ENGINE_MULTITHREADED, **kw):
    """the same case on the reference CPU engine and on Engine_CUDA (both inside the reference harness)"""
    with backend(lambda *args: RefSim(*args, engine=cpu_engine, threads=3, cuda_lib=True)):
        c = fn(*a, **kw)
    with backend(lambda *args: RefSim(*args, engine=ENGINE_CUDA, fast_processing=fast)):
        g = fn(*a, **kw)
    c = c[0] if isinstance(c, tuple) else c
    g = g[0] if isinstance(g, tuple) else g
    return c, g


def fields_equal(c, g, what):
    assert c.num_ts == g.num_ts
    assert_same(c.volt, g.volt, what + " volt")
    assert_same(c.curr, g.curr, what + " curr")


@pytest.mark.parametrize("case", ["cavity", "allpml", "c1_sinus", "c4_drude", "c3_patch"])
def test_engine_cuda_equals_reference_engine_fields(case):
    fn, args, kw, steps = {
        "cavity": (cases.engine_cavity, (), {}, (1, 2, 150)),
        "allpml": (cases.uniform_box, (), dict(n=(29, 23, 31)), (1, 60)),
        "c1_sinus": (configs.c1_parallel_plate_waveguide, ("sinus",), {}, (1, 120)),
        "c4_drude": (configs.c4_drude_block, (), dict(n=(30, 30, 30), block=(10, 20)), (1, 60)),
        "c3_patch": (configs.c3_patch_antenna, (), dict(n=(40, 40, 30)), (40,)),
    }[case]
    c, g = pair(fn, *args, **kw)
    assert_operator_equal(c, g, case)   # same reference operator code on both sides
    for n in steps:
        c.iterate(n)
        g.iterate(n)
        fields_equal(c, g, "%s @%d" % (case, c.num_ts))
    assert np.abs(c.volt).max() > 0 and np.abs(c.curr).max() > 0
    # read-outs through Engine_Interface_CUDA_FDTD
    N = c.N
    a, b = (N[0] // 2, N[1] // 2, 2), (N[0] // 2, N[1] // 2, N[2] - 3)
    assert c.voltage_integral(a, b) == g.voltage_integral(a, b)
    p = (N[0] // 2, N[1] // 2, N[2] // 2)
    for h in (0, 1):
        assert np.array_equal(c.raw_field(h, p), g.raw_field(h, p))
    # CalcFastEnergy: the device evaluates Engine_Interface_FDTD::CalcFastEnergy (engine_interface_fdtd.cpp:302-350,
    # lines 0..N-2 of every direction, fp64 sums).  Engine_Interface_SSE_FDTD::CalcFastEnergy
    # (engine_interface_sse_fdtd.cpp:40-75) sums float lanes over ALL z vectors -- it includes the last z line, which
    # carries tangential E on a Mur face (C1: +1.6 %) -- so the reference value is bracketed by the two sums of the
    # (bit-equal) fields instead of being compared directly
    ec, eg = c.energy(), g.energy()
    v, i = np.asarray(c.volt, np.float32), np.asarray(c.curr, np.float32)
    def esum(zend):   # float products, fp64 sums (FDTD_FLOAT * FDTD_FLOAT added to a double)
        return 8.85418781762e-12 * (v[:, :-1, :-1, :zend] * v[:, :-1, :-1, :zend]).sum(dtype=np.float64) \
            + 1.256637062e-6 * (i[:, :-1, :-1, :zend] * i[:, :-1, :-1, :zend]).sum(dtype=np.float64)
    assert eg > 0 and abs(eg - esum(-1)) <= 1e-11 * eg
    assert abs(ec - esum(None)) <= 1e-4 * ec


def _probe_setup(s, lines):
    vbox = ((5, 4, 5), (9, 4, 5))
    cbox = ((4, 3, 10), (12, 8, 10))
    fpos = (16, 6, 20)
    s.add_probe(0, "ut1", [lines[n][vbox[0][n]] for n in range(3)], [lines[n][vbox[1][n]] for n in range(3)])
    c0 = [s.disc_line(n, cbox[0][n], True) for n in range(3)]
    c1 = [s.disc_line(n, cbox[1][n], True) for n in range(3)]
    s.add_probe(1, "it1", c0, c1, norm_dir=2)
    fp = [lines[n][fpos[n]] for n in range(3)]
    s.add_probe(2, "et1", fp, fp)
    s.add_probe(3, "ht1", fp, fp)


@pytest.mark.parametrize("fast", [True, False], ids=["Process_CUDA-classes", "stock-Processing-classes"])
def test_reference_processing_classes_write_identical_probe_files(tmp_path, fast):
    """ProcessVoltage / ProcessCurrent / ProcessFieldProbe, run by the RunFDTD loop: the ASCII files of the CUDA run are
    byte-identical to those of the reference's multithreaded engine (port voltage/current series: rel-L2 = 0)"""
    cwd = os.getcwd()
    out = {}
    try:
        c, g = pair(cases.engine_cavity, fast=fast)
        lines = (c.x, c.y, c.z)
        for tag, s in (("cpu", c), ("gpu", g)):
            d = tmp_path / tag
            d.mkdir()
            os.chdir(d)
            _probe_setup(s, lines)
            s.run(240)
            out[tag] = {f: pyref.read_probe_file(f) for f in ("ut1", "it1", "et1", "ht1")}
        for f in out["cpu"]:
            a, b = out["cpu"][f], out["gpu"][f]
            assert a.shape == b.shape and a.shape[0] > 20
            assert np.array_equal(a, b), f          # identical at the 12 printed digits
            assert np.abs(a[:, 1:]).max() > 0
        fields_equal(c, g, "after RunFDTD")
        assert np.abs(c.volt).max() > 0
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("fast", [True, False], ids=["Process_CUDA-classes", "stock-Processing-classes"])
def test_reference_field_dumps_td_fd_and_mode_match(tmp_path, fast):
    """ProcessFieldsTD (HDF5 + VTK, cell / node / no interpolation, E and H), ProcessFieldsFD (running DFT) and
    ProcessModeMatch of the reference on both engines: recorded datasets equal bit for bit"""
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        def case():
            x = np.arange(24) * 1e-3
            y = np.arange(16) * 1e-3
            z = np.arange(60) * 1e-3
            s = cases.OracleSim(x, y, z, 1.0)
            s.set_bc([BC_PEC, BC_PEC, BC_PEC, BC_PEC, BC_PML, BC_PML], (6,) * 6)
            s.set_excite_gauss(9e9, 3e9)
            s.add_excitation((x[0], y[0], z[12]), (x[-1], y[-1], z[12]), EXC_E_SOFT, (0, 1, 0))
            s.build()
            return s
        c, g = pair(case, fast=fast)
        lo, hi = (0.004, 0.003, 0.015), (0.018, 0.012, 0.040)
        res = {}
        for tag, s, cuda in (("cpu", c, True), ("gpu", g, True)):
            pyref.recorded_clear(cuda)
            s.add_dump("Et_cell", lo, hi, dump_type=0, file_type=1, interp=2, interval=20)
            s.add_dump("Ht_cell", lo, hi, dump_type=1, file_type=1, interp=2, interval=20)
            s.add_dump("Et_node_vtk", lo, hi, dump_type=0, file_type=0, interp=1, interval=40)
            s.add_dump("Ht_raw", (0.004, 0.003, 0.030), (0.018, 0.012, 0.030), dump_type=1, file_type=1, interp=0, interval=20)
            s.add_fd_dump("Ef", lo, hi, [7e9, 9e9, 11e9], dump_type=0, interp=2)
            s.add_fd_dump("Hf", lo, hi, [9e9], dump_type=1, interp=2)
            s.add_mode_match("mm_e", (0.0, 0.0, 0.040), (0.023, 0.015, 0.040), 0, "0", "sin(pi*x/0.023)", 2)
            s.add_mode_match("mm_h", (0.0, 0.0, 0.040), (0.023, 0.015, 0.040), 1, "-sin(pi*x/0.023)", "0", 2)
            s.run(300)
            rec = pyref.recorded(cuda)
            res[tag] = (rec, pyref.read_probe_file("mm_e"), pyref.read_probe_file("mm_h"))
            for f in ("mm_e", "mm_h"):
                os.rename(f, f + "." + tag)
        rc, rg = res["cpu"][0], res["gpu"][0]
        assert sorted(rc) == sorted(rg) and len(rc) > 30
        n_td = 0
        for k in rc:
            assert rc[k].shape == rg[k].shape, k
            if "@" in k and "time" in k:
                assert np.array_equal(rc[k], rg[k]), k
                continue
            assert np.array_equal(rc[k].astype(np.float32).view(np.uint32), rg[k].astype(np.float32).view(np.uint32)) \
                or np.array_equal(rc[k], rg[k]), k
            n_td += "/FieldData/TD/" in k
        assert n_td >= 20
        assert max(np.abs(v).max() for k, v in rc.items() if "/FieldData/FD/f" in k) > 0
        for i in (1, 2):
            assert np.array_equal(res["cpu"][i], res["gpu"][i])
            assert np.abs(res["cpu"][i][:, 1]).max() > 0
        fields_equal(c, g, "after dumps")
    finally:
        os.chdir(cwd)


def test_conducting_sheet_and_lumped_rlc_built_by_the_reference():
    """Operator_Ext_ConductingSheet (operator_ext_conductingsheet.h:31, a Lorentz-type ADE list) and the
    series/parallel lumped RLC builder (operator_ext_lumpedRLC.cpp:112-534): tables built by the reference,
    uploaded by Engine_CUDA::InitExtensions, fields equal to the reference engine's"""
    def case():
        lines = tuple(np.arange(m) * 1e-3 for m in (26, 24, 28))
        s = cases.OracleSim(*lines, 1.0)
        s.set_bc([BC_MUR, BC_MUR, BC_PML, BC_PML, BC_PEC, BC_MUR], (6,) * 6)
        s.set_excite_gauss(5e9, 5e9)
        s.add_conducting_sheet((0.006, 0.006, 0.012), (0.018, 0.016, 0.012), 56e6, 18e-6)
        s.add_lumped_rlc((0.010, 0.010, 0.004), (0.010, 0.010, 0.008), 2, R=50.0, Cap=1e-12, L=1e-9, series=True)
        s.add_lumped_rlc((0.014, 0.012, 0.004), (0.014, 0.012, 0.008), 2, R=75.0, Cap=0.5e-12, L=2e-9, series=False)
        c = cases.edge_center(lines, 2, (12, 11, 6))
        s.add_excitation(c, c, EXC_E_SOFT, (0, 0, 1))
        s.build()
        return s
    c, g = pair(case)
    assert len(c.lorentz_extensions()) >= 1 and c.lorentz_extensions()[0][0]["count"] > 50
    assert c.rlc_tables()[0].size >= 8
    for n in (1, 2, 3, 100):
        c.iterate(n)
        g.iterate(n)
        fields_equal(c, g, "sheet+rlc @%d" % c.num_ts)
