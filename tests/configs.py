"""BASELINE.json configs C1-C4 as CSXCAD-free box geometries.  The MSL and patch
meshes are hand-built uniform-Delta equivalents of the tutorials (the tutorials' own meshes
need CSXCAD's SmoothMeshLines).  Each builder returns a built OracleSim plus the probe set."""
import numpy as np

from oracle.pyoracle import OracleSim, BC_PEC, BC_PMC, BC_MUR, BC_PML, EXC_E_SOFT

C0 = 299792458.0


def c1_parallel_plate_waveguide(excite="gauss"):
    """matlab/Tutorials/Parallel_Plate_Waveguide.m:19-37: 21x21x41 nodes, unit 1 m,
    BC {PMC,PMC,PEC,PEC,MUR,MUR}, E_y soft source on the plane z=0"""
    x = np.arange(-10, 11, dtype=np.float64)
    y = np.arange(-10, 11, dtype=np.float64)
    z = np.arange(-10, 31, dtype=np.float64)
    s = OracleSim(x, y, z, 1.0)
    s.set_bc([BC_PMC, BC_PMC, BC_PEC, BC_PEC, BC_MUR, BC_MUR])
    if excite == "gauss":
        s.set_excite_gauss(5e6, 5e6)
    else:
        s.set_excite_sinus(10e6)
    s.add_excitation((-10, -10, 0), (10, 10, 0), EXC_E_SOFT, (0, 1, 0))
    s.build()
    probes = dict(volt=[((10, 0, 25), (10, 20, 25)), ((10, 0, 35), (10, 20, 35))],
                  curr=[(((0, 0, 25), (20, 20, 25)), 2, (1, 1, 1), (1, 1, 1))],
                  field=[(0, (10, 10, 30)), (1, (10, 10, 30))])
    return s, probes


def c2_msl_notch_filter(n=(120, 81, 21)):
    """python/Tutorials/MSL_NotchFilter.py:28-92 on a uniform mesh: BC {PML_8,PML_8,MUR,MUR,PEC,MUR},
    Gauss f0=fc=3.5 GHz, substrate eps_r 3.66, a through line with an open stub, two MSL ports
    (3 voltage + 2 current probes each, python/openEMS/ports.py:266-294)"""
    nx, ny, nz = n
    x = np.arange(nx) * 0.5    # mm
    y = np.arange(ny) * 0.5
    z = np.arange(nz) * 0.254
    s = OracleSim(x, y, z, 1e-3)
    s.set_bc([BC_PML, BC_PML, BC_MUR, BC_MUR, BC_PEC, BC_MUR])
    s.set_excite_gauss(3.5e9, 3.5e9)
    hs = 4                      # substrate: 4 cells = 1.016 mm
    jc = ny // 2
    s.add_material((x[0], y[0], z[0]), (x[-1], y[-1], z[hs]), epsR=3.66)
    s.add_metal((x[0], y[jc - 2], z[hs]), (x[-1], y[jc + 2], z[hs]))          # through line, 2 mm wide
    s.add_metal((x[nx // 2 - 2], y[jc + 2], z[hs]), (x[nx // 2 + 2], y[jc + 26], z[hs]))  # open stub 12 mm
    p1, p2 = 20, nx - 21        # port planes (inside the PML-free region)
    s.add_excitation((x[p1 - 6], y[jc - 2], z[0]), (x[p1 - 6], y[jc + 2], z[hs]), EXC_E_SOFT, (0, 0, 1))
    s.build()
    ports = []
    for p, direction in ((p1, 1), (p2, -1)):
        u = [((p + d, jc, 0), (p + d, jc, hs)) for d in (-1, 0, 1)]
        i = [(((p + d, jc - 4, hs - 2), (p + d, jc + 4, hs + 2)), 0, (1, 1, 1), (1, 1, 1)) for d in (-1, 0)]
        ports.append(dict(volt=u, curr=i, direction=direction, delta=0.5e-3))
    return s, ports


def msl_port_spectra(port, ut, it, t_u, t_i, freq):
    """python/openEMS/ports.py:319-339 (MSLPort.ReadUIData) + :115-151 (CalcPort)"""
    def dft(t, v):
        dt = t[1] - t[0]
        return np.array([np.sum(v * np.exp(-2j * np.pi * f * t)) * dt * 2 for f in freq])
    U = [dft(t_u, u) for u in ut]
    I = [dft(t_i, i) * port["direction"] for i in it]
    uf_tot = U[1]
    if_tot = 0.5 * (I[0] + I[1])
    Et, dEt = U[1], (U[2] - U[0]) / (2 * port["delta"])
    Ht, dHt = if_tot, (I[1] - I[0]) / port["delta"]
    Z = np.sqrt(Et * dEt / (Ht * dHt))
    uf_inc = 0.5 * (uf_tot + if_tot * Z)
    uf_ref = uf_tot - uf_inc
    return uf_inc, uf_ref, Z


def c3_patch_antenna(n=(70, 70, 40)):
    """python/Tutorials/Simple_Patch_Antenna.py with PML_8 on all faces (BASELINE C3): substrate,
    ground, patch, lumped 50 Ohm port (parallel RC folded into vv/vi, operator.cpp:1586-1763),
    NF2FF box = E and H cell-interpolated dumps on 6 faces (python/openEMS/nf2ff.py:61-96)"""
    nx, ny, nz = n
    x = (np.arange(nx) - nx // 2) * 2.0   # mm
    y = (np.arange(ny) - ny // 2) * 2.0
    z = (np.arange(nz) - 12) * 0.762
    s = OracleSim(x, y, z, 1e-3)
    s.set_bc([BC_PML] * 6, (8,) * 6)
    s.set_excite_gauss(2e9, 1e9)
    hs = 2                                 # substrate 1.524 mm = 2 cells above z=0
    k0 = 12
    s.add_material((-30, -30, 0), (30, 30, z[k0 + hs]), epsR=3.38, kappa=1e-3 * 2 * np.pi * 2.45e9 * 8.85418781762e-12 * 3.38)
    s.add_metal((-30, -30, 0), (30, 30, 0))                       # ground
    s.add_metal((-16, -20, z[k0 + hs]), (16, 20, z[k0 + hs]))     # patch 32 x 40 mm
    feed = (-6.0, 0.0)
    s.add_lumped_rc((feed[0], feed[1], 0), (feed[0], feed[1], z[k0 + hs]), 2, R=50.0, caps=True)
    s.add_excitation((feed[0], feed[1], 0), (feed[0], feed[1], z[k0 + hs]), EXC_E_SOFT, (0, 0, 1))
    s.build()
    lo = (10, 10, 4)
    hi = (nx - 11, ny - 11, nz - 11)
    faces = []
    for a in range(3):
        for side in (lo, hi):
            start, stop = list(lo), list(hi)
            start[a] = stop[a] = side[a]
            faces.append((tuple(start), tuple(stop)))
    fi = int(np.argmin(np.abs(x - feed[0]))), int(np.argmin(np.abs(y - feed[1])))
    port = dict(volt=[((fi[0], fi[1], k0), (fi[0], fi[1], k0 + hs))],
                curr=[(((fi[0] - 1, fi[1] - 1, k0 + 1), (fi[0], fi[1], k0 + 1)), 2, (1, 1, 1), (1, 1, 1))])
    return s, port, faces


def c4_drude_block(n=(48, 48, 48), block=(12, 36)):
    """BASELINE C4 in small: uniform mesh, PML_8 x6, central Drude eps+mue block (f_p 5 GHz,
    tau 5 ns: matlab/examples/other/Metamaterial_PlaneWave_Drude.m:28-33,75-76), plane E_y source"""
    lines = tuple(np.arange(m, dtype=np.float64) for m in n)
    s = OracleSim(*lines, 1e-3)
    s.set_bc([BC_PML] * 6, (8,) * 6)
    s.set_excite_gauss(5e9, 5e9)
    a, b = block
    s.add_lorentz((a, a, a), (b, b, b), eps_fp=(5e9,), eps_tau=(5e-9,), mue_fp=(5e9,), mue_tau=(5e-9,))
    s.add_excitation((10, 10, 10), (n[0] - 11, n[1] - 11, 10), EXC_E_SOFT, (0, 1, 0))
    s.build()
    return s
