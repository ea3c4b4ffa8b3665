"""Shared problem definitions for the parity tests (used with the oracle AND the CUDA engine).

Each builder returns an oracle.pyoracle.OracleSim already built; tests/ then hand its
operator to the CUDA engine through the C-ABI (openems_b200) and compare bit for bit.
"""
import numpy as np

from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT

C0 = 299792458.0


def edge_center(lines, n, pos):
    c = [lines[0][pos[0]], lines[1][pos[1]], lines[2][pos[2]]]
    c[n] = 0.5 * (lines[n][pos[n]] + lines[n][pos[n] + 1])
    return c


def analytic_cavity(max_ts=20000):
    """TESTSUITE/combinedtests/cavity.m:37-107: PEC cavity 5x2x6 cm, mesh 26x11x32, Gauss
    f0=fc=4.5 GHz, curve excitation between two diagonal neighbour nodes (path z,x,y found by
    Operator::FindPath operator.cpp:390-480 for this mesh)."""
    a, b, d = 5e-2, 2e-2, 6e-2
    lines = (np.linspace(0, a, 26), np.linspace(0, b, 11), np.linspace(0, d, 32))
    s = OracleSim(*lines, 1.0)
    s.set_bc([BC_PEC] * 6)
    s.set_excite_gauss(4.5e9, 4.5e9)
    for n, pos in ((2, (16, 6, 20)), (0, (16, 6, 21)), (1, (17, 6, 21))):
        c = edge_center(lines, n, pos)
        s.add_excitation(c, c, EXC_E_SOFT, (1, 1, 1))
    s.build(max_ts)
    probes = dict(ut1x=((5, 4, 5), (6, 4, 5)), ut1y=((5, 4, 5), (5, 5, 5)), ut1z=((12, 4, 5), (12, 4, 6)))
    return s, probes, (a, b, d)


def engine_cavity(max_ts=1000, n=(27, 11, 33), bc=(BC_MUR, BC_PML, BC_PMC, BC_PEC, BC_PEC, BC_PEC)):
    """TESTSUITE/enginetests/cavity.m:70-115: mesh 27x11x33 over 5x2x6 cm, BC
    {MUR, PML_8, PMC, PEC, PEC, PEC}, Gauss excite, a dielectric box and an excited curve."""
    a, b, d = 5e-2, 2e-2, 6e-2
    lines = (np.linspace(0, a, n[0]), np.linspace(0, b, n[1]), np.linspace(0, d, n[2]))
    s = OracleSim(*lines, 1.0)
    s.set_bc(list(bc))
    s.set_excite_gauss(4.5e9, 4.5e9)
    # dielectric box in the lower third
    s.add_material((lines[0][3], lines[1][2], lines[2][4]), (lines[0][9], lines[1][6], lines[2][12]),
                   epsR=3.5, kappa=0.02)
    i, j, k = (2 * n[0]) // 3, (2 * n[1]) // 3, (2 * n[2]) // 3
    for c_n, pos in ((2, (i, j, k)), (0, (i, j, k + 1)), (1, (i + 1, j, k + 1))):
        c = edge_center(lines, c_n, pos)
        s.add_excitation(c, c, EXC_E_SOFT, (1, 1, 1))
    s.build(max_ts)
    return s


def uniform_box(n=(40, 36, 44), delta=1e-3, bc=(BC_PML,) * 6, pml=8, f0=0.0, fc=None, max_ts=10 ** 6,
                materials=(), metals=(), lorentz=(), src_comp=2, src_pos=None, extra=None):
    """synthetic uniform Cartesian mesh (BASELINE config C5 in small): vacuum, Delta, point E source"""
    lines = tuple(np.arange(m) * delta for m in n)
    s = OracleSim(*lines, 1.0)
    s.set_bc(list(bc), (pml,) * 6)
    if fc is None:
        fc = C0 / (20 * delta)
    s.set_excite_gauss(f0, fc)
    for m in materials:
        s.add_material(**m)
    for m in metals:
        s.add_metal(**m)
    for m in lorentz:
        s.add_lorentz(**m)
    if src_pos is None:
        src_pos = tuple(m // 2 for m in n)
    c = edge_center(lines, src_comp, src_pos)
    vec = [0, 0, 0]
    vec[src_comp] = 1
    s.add_excitation(c, c, EXC_E_SOFT, vec)
    if extra:
        extra(s, lines)
    s.build(max_ts)
    return s
