"""ctypes binding of the CPU oracle (oracle/fdtd_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under openems_b200/ does.  Arrays handed out are numpy views in
the reference's ArrayNIJK order [n][i][j][k] (k fastest, tools/arraylib/array_nijk.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

BC_PEC, BC_PMC, BC_MUR, BC_PML = 0, 1, 2, 3
EXC_E_SOFT, EXC_E_HARD, EXC_H_SOFT, EXC_H_HARD = 0, 1, 2, 3

_u3 = C.c_uint * 3
_d3 = C.c_double * 3
_i3 = C.c_int * 3
_i6 = C.c_int * 6
_u6 = C.c_uint * 6
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint)
_dp = C.POINTER(C.c_double)


def build(force=False):
    """compile liboracle.so with oracle/Makefile (building the checker is not using it)"""
    srcs = [os.path.join(_HERE, f) for f in ("fdtd_oracle.c", "fdtd_oracle_sse.c", "fdtd_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp = C.c_void_p
    sig = {
        "orc_create": (vp, [_up, _dp, _dp, _dp, C.c_double]),
        "orc_destroy": (None, [vp]),
        "orc_set_bc": (None, [vp, _i6, _u6]),
        "orc_set_background": (None, [vp] + [C.c_double] * 4),
        "orc_set_mur_phase_velocity": (None, [vp, C.c_double]),
        "orc_set_timestep": (None, [vp, C.c_double, C.c_double]),
        "orc_add_material": (C.c_int, [vp, C.c_int, _d3, _d3] + [C.c_double] * 4),
        "orc_add_metal": (C.c_int, [vp, C.c_int, _d3, _d3]),
        "orc_add_lorentz": (C.c_int, [vp, C.c_int, _d3, _d3] + [C.c_double] * 4 + [C.c_int] + [_dp] * 6),
        "orc_add_excitation": (C.c_int, [vp, C.c_int, _d3, _d3, C.c_int, _d3, C.c_double]),
        "orc_add_lumped_rc": (C.c_int, [vp, _d3, _d3, C.c_int, C.c_double, C.c_double, C.c_int]),
        "orc_add_rlc_raw": (C.c_int, [vp, C.c_uint, C.POINTER(C.c_int), _up] + [_fp] * 9),
        "orc_add_steadystate": (C.c_int, [vp, C.c_uint, C.c_uint, _up, C.POINTER(C.c_int)]),
        "orc_set_tfsf": (C.c_int, [vp, _u3, _u3, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "orc_tfsf_on": (C.c_int, [vp]),
        "orc_tfsf_max_delay": (C.c_uint, [vp]),
        "orc_tfsf_box": (None, [vp, _u3, _u3, C.POINTER(C.c_int)]),
        "orc_tfsf_face": (C.c_uint, [vp, C.c_int, C.c_int, C.c_int, C.c_int, _up, _fp, _fp]),
        "orc_add_absorbing_sheet": (C.c_int, [vp, _u3, _u3, C.c_int, C.c_int, C.c_double]),
        "orc_abc_count": (C.c_int, [vp]),
        "orc_abc_info": (None, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _u3, _u3]),
        "orc_abc_coeff": (None, [vp, C.c_int, _fp, _fp, _fp, _fp]),
        "orc_steadystate_last_diff": (C.c_double, [vp]),
        "orc_set_excite_gauss": (None, [vp, C.c_double, C.c_double]),
        "orc_set_excite_sinus": (None, [vp, C.c_double]),
        "orc_set_excite_dirac": (None, [vp, C.c_double]),
        "orc_set_excite_step": (None, [vp, C.c_double]),
        "orc_build": (C.c_int, [vp, C.c_uint]),
        "orc_dT": (C.c_double, [vp]),
        "orc_nyquist": (C.c_uint, [vp]),
        "orc_coeff": (_fp, [vp, C.c_int]),
        "orc_signal_length": (C.c_uint, [vp]),
        "orc_signal": (_fp, [vp, C.c_int]),
        "orc_signal_period_ts": (C.c_uint, [vp]),
        "orc_exc_count": (C.c_uint, [vp, C.c_int]),
        "orc_exc_get": (None, [vp, C.c_int, _up, _up, _fp, _up]),
        "orc_upml_count": (C.c_int, [vp]),
        "orc_upml_box": (None, [vp, C.c_int, _u3, _u3]),
        "orc_upml_coeff": (_fp, [vp, C.c_int, C.c_int]),
        "orc_upml_flux": (_fp, [vp, C.c_int, C.c_int]),
        "orc_mur_count": (C.c_int, [vp]),
        "orc_mur_info": (None, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _up, _up, C.c_uint * 2, _up]),
        "orc_mur_coeff": (_fp, [vp, C.c_int, C.c_int]),
        "orc_lorentz_order": (C.c_int, [vp]),
        "orc_lorentz_count": (C.c_uint, [vp, C.c_int]),
        "orc_lorentz_flags": (C.c_int, [vp, C.c_int]),
        "orc_lorentz_pos": (_up, [vp, C.c_int, C.c_int]),
        "orc_lorentz_coeff": (_fp, [vp, C.c_int, C.c_int, C.c_int]),
        "orc_iterate": (None, [vp, C.c_uint]),
        "orc_num_ts": (C.c_uint, [vp]),
        "orc_volt": (_fp, [vp]),
        "orc_curr": (_fp, [vp]),
        "orc_reset_fields": (None, [vp]),
        "orc_voltage_integral": (C.c_double, [vp, _u3, _u3]),
        "orc_current_integral": (C.c_double, [vp, _u3, _u3, C.c_int, _i3, _i3]),
        "orc_raw_field": (None, [vp, C.c_int, _u3, _d3]),
        "orc_energy": (C.c_double, [vp]),
        "orc_dump_field": (None, [vp, C.c_int, C.c_int, _u3, _u3, _fp]),
        "orc_mode_match": (None, [vp, C.c_int, C.c_int, _u3, _u3, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "orc_fd_weight": (None, [C.c_double, C.c_double, C.c_double, C.c_uint, _fp]),
        "orc_fd_accumulate": (None, [_fp, _fp, C.c_size_t, _fp]),
        "orc_edge_length": (C.c_double, [vp, C.c_int, _u3, C.c_int]),
        "orc_disc_line": (C.c_double, [vp, C.c_int, C.c_uint, C.c_int]),
        # sse-compressed multithreaded restatement (fdtd_oracle_sse.c)
        "orc_sse_create": (vp, [vp, C.c_int]),
        "orc_sse_destroy": (None, [vp]),
        "orc_sse_iterate": (None, [vp, C.c_uint]),
        "orc_sse_unique": (C.c_uint, [vp]),
        "orc_sse_get_fields": (None, [vp, _fp, _fp]),
        "orc_sse_num_ts": (C.c_uint, [vp]),
    }
    SIGNATURES.update(sig)
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _darr(v):
    return np.ascontiguousarray(v, dtype=np.float64)


class _Namespace:
    """attribute access to prefix+name of a ctypes library"""

    def __init__(self, L, prefix):
        self._L, self._prefix = L, prefix

    def __getattr__(self, name):
        return getattr(self._L, self._prefix + name)


SIGNATURES = {}


class OracleSim:
    """One reference-semantics FDTD setup + scalar engine (Engine, FDTD/engine.cpp)."""

    def _functions(self):
        """namespace of the C functions without their prefix (orc_*; oracle/pyref.py serves ref_* here)"""
        return _Namespace(lib(), "orc_")

    def __init__(self, x, y, z, grid_delta=1.0):
        self._f = self._functions()
        self.x, self.y, self.z = _darr(x), _darr(y), _darr(z)
        self.N = (len(self.x), len(self.y), len(self.z))
        nl = _u3(*self.N)
        self._h = self._f.create(nl, self.x.ctypes.data_as(_dp), self.y.ctypes.data_as(_dp),
                               self.z.ctypes.data_as(_dp), grid_delta)
        if not self._h:
            raise ValueError("oracle: need at least 3 lines per direction")
        self.grid_delta = grid_delta
        self._keep = []

    def close(self):
        if self._h:
            self._f.destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def set_bc(self, bc, pml_size=(8,) * 6):
        self._f.set_bc(self._h, _i6(*bc), _u6(*pml_size))

    def set_background(self, epsR=1.0, mueR=1.0, kappa=0.0, sigma=0.0):
        self._f.set_background(self._h, epsR, mueR, kappa, sigma)

    def set_timestep(self, forced_dT=0.0, factor=1.0):
        self._f.set_timestep(self._h, forced_dT, factor)

    def set_mur_phase_velocity(self, v):
        self._f.set_mur_phase_velocity(self._h, v)

    def add_material(self, start, stop, epsR=1.0, mueR=1.0, kappa=0.0, sigma=0.0, prio=0):
        return self._f.add_material(self._h, prio, _d3(*start), _d3(*stop), epsR, mueR, kappa, sigma)

    def add_metal(self, start, stop, prio=10):
        return self._f.add_metal(self._h, prio, _d3(*start), _d3(*stop))

    def add_lorentz(self, start, stop, epsR=1.0, mueR=1.0, kappa=0.0, sigma=0.0, prio=0,
                    eps_fp=(), eps_tau=(), eps_flor=(), mue_fp=(), mue_tau=(), mue_flor=()):
        order = max(len(eps_fp), len(mue_fp))

        def arr(v):
            a = np.zeros(order, dtype=np.float64)
            a[:len(v)] = v
            return a
        arrs = [arr(v) for v in (eps_fp, eps_tau, eps_flor, mue_fp, mue_tau, mue_flor)]
        return self._f.add_lorentz(self._h, prio, _d3(*start), _d3(*stop), epsR, mueR, kappa, sigma,
                                     order, *[a.ctypes.data_as(_dp) for a in arrs])

    def add_excitation(self, start, stop, exc_type, vec, delay=0.0, prio=0):
        return self._f.add_excitation(self._h, prio, _d3(*start), _d3(*stop), exc_type, _d3(*vec), delay)

    def add_lumped_rc(self, start, stop, direction, R=float("nan"), Cap=float("nan"), caps=True):
        return self._f.add_lumped_rc(self._h, _d3(*start), _d3(*stop), direction, R, Cap, int(caps))

    def add_rlc_raw(self, direction, pos, coeffs):
        """direction: int[count]; pos: uint[3][count]; coeffs: dict of the 9 arrays"""
        d = np.ascontiguousarray(direction, dtype=np.int32)
        p = np.ascontiguousarray(pos, dtype=np.uint32)
        names = ("ilv", "i2v", "vvd", "vv2", "vj1", "vj2", "ib0", "b1", "b2")
        arrs = [np.ascontiguousarray(coeffs[k], dtype=np.float32) for k in names]
        return self._f.add_rlc_raw(self._h, len(d), d.ctypes.data_as(C.POINTER(C.c_int)),
                                     p.ctypes.data_as(_up), *[a.ctypes.data_as(_fp) for a in arrs])

    def add_steadystate(self, period_ts, pos3, direction):
        p = np.ascontiguousarray(pos3, dtype=np.uint32)
        d = np.ascontiguousarray(direction, dtype=np.int32)
        return self._f.add_steadystate(self._h, int(period_ts), len(d), p.ctypes.data_as(_up), d.ctypes.data_as(C.POINTER(C.c_int)))

    def steadystate_last_diff(self):
        return self._f.steadystate_last_diff(self._h)

    def set_excite_gauss(self, f0, fc):
        self._f.set_excite_gauss(self._h, f0, fc)

    def set_excite_sinus(self, f0):
        self._f.set_excite_sinus(self._h, f0)

    def set_excite_dirac(self, fmax):
        self._f.set_excite_dirac(self._h, fmax)

    def set_excite_step(self, fmax):
        self._f.set_excite_step(self._h, fmax)

    def build(self, max_ts=10 ** 9):
        rc = self._f.build(self._h, max_ts)
        if rc != 0:
            raise RuntimeError("oracle build failed rc=%d" % rc)

    # ---- operator results
    @property
    def dT(self):
        return self._f.dT(self._h)

    @property
    def nyquist(self):
        return self._f.nyquist(self._h)

    def _field_view(self, ptr):
        n = 3 * self.N[0] * self.N[1] * self.N[2]
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(3, *self.N)

    def coeff(self, which):
        """'vv' | 'vi' | 'ii' | 'iv' -> view [3][Nx][Ny][Nz]"""
        return self._field_view(self._f.coeff(self._h, ("vv", "vi", "ii", "iv").index(which)))

    def signal(self):
        n = self._f.signal_length(self._h)
        v = np.ctypeslib.as_array(self._f.signal(self._h, 0), shape=(n,)).copy()
        i = np.ctypeslib.as_array(self._f.signal(self._h, 1), shape=(n,)).copy()
        return v, i, self._f.signal_period_ts(self._h)

    def excitation(self, is_curr):
        n = self._f.exc_count(self._h, int(is_curr))
        idx = np.zeros((3, n), dtype=np.uint32)
        d = np.zeros(n, dtype=np.uint32)
        amp = np.zeros(n, dtype=np.float32)
        delay = np.zeros(n, dtype=np.uint32)
        if n:
            self._f.exc_get(self._h, int(is_curr), idx.ctypes.data_as(_up), d.ctypes.data_as(_up),
                              amp.ctypes.data_as(_fp), delay.ctypes.data_as(_up))
        return idx, d, amp, delay

    def upml_boxes(self):
        out = []
        for b in range(self._f.upml_count(self._h)):
            st, nl = _u3(), _u3()
            self._f.upml_box(self._h, b, st, nl)
            shape = (3, nl[0], nl[1], nl[2])
            cnt = int(np.prod(shape))
            co = [np.ctypeslib.as_array(self._f.upml_coeff(self._h, b, w), shape=(cnt,)).reshape(shape)
                  for w in range(6)]
            out.append(dict(start=tuple(st), n=tuple(nl), vv=co[0], vvfn=co[1], vvfo=co[2],
                            ii=co[3], iifn=co[4], iifo=co[5]))
        return out

    def upml_flux(self, b, is_curr):
        st, nl = _u3(), _u3()
        self._f.upml_box(self._h, b, st, nl)
        shape = (3, nl[0], nl[1], nl[2])
        return np.ctypeslib.as_array(self._f.upml_flux(self._h, b, int(is_curr)),
                                     shape=(int(np.prod(shape)),)).reshape(shape)

    def set_tfsf(self, start, stop, prop_dir, e_amp):
        """plane-wave (TFSF) excitation on the box of mesh indices start..stop"""
        pd = (C.c_double * 3)(*prop_dir)
        ea = (C.c_double * 3)(*e_amp)
        rc = self._f.set_tfsf(self._h, _u3(*start), _u3(*stop), pd, ea)
        if rc:
            raise RuntimeError("orc_set_tfsf rc=%d" % rc)

    def tfsf(self):
        """tables of Operator_Ext_TFSF after build: dict or None"""
        if not self._f.tfsf_on(self._h):
            return None
        start, stop = (C.c_uint * 3)(), (C.c_uint * 3)()
        act = (C.c_int * 6)()
        self._f.tfsf_box(self._h, start, stop, act)
        nl = [stop[n] - start[n] + 1 for n in range(3)]
        faces = {}
        for which in (0, 1):
            for n in range(3):
                numP = nl[(n + 1) % 3] * nl[(n + 2) % 3]
                for l in range(2):
                    if not act[2 * n + l]:
                        continue
                    for c in range(2):
                        d = np.zeros(numP, np.uint32); dd = np.zeros(numP, np.float32); a = np.zeros(numP, np.float32)
                        self._f.tfsf_face(self._h, which, n, l, c, d.ctypes.data_as(_up), dd.ctypes.data_as(_fp), a.ctypes.data_as(_fp))
                        faces[(which, n, l, c)] = (d, dd, a)
        return dict(start=tuple(start), stop=tuple(stop), active=[[act[2 * n], act[2 * n + 1]] for n in range(3)],
                    max_delay=self._f.tfsf_max_delay(self._h), faces=faces)

    def add_absorbing_sheet(self, x0, x1, normal_positive, abc_type, phase_velocity=0.0):
        """local absorbing sheet on mesh indices x0..x1 (one direction single-line); type 1 Mur 1st order,
        2 with super-absorption"""
        rc = self._f.add_absorbing_sheet(self._h, _u3(*x0), _u3(*x1), int(normal_positive), int(abc_type), float(phase_velocity))
        if rc:
            raise RuntimeError("orc_add_absorbing_sheet rc=%d" % rc)

    def absorbing_sheets(self):
        out = []
        for a in range(self._f.abc_count(self._h)):
            ny, ty, pos = C.c_int(), C.c_int(), C.c_int()
            x0, x1 = (C.c_uint * 3)(), (C.c_uint * 3)()
            self._f.abc_info(self._h, a, C.byref(ny), C.byref(ty), C.byref(pos), x0, x1)
            nP, nPP = (ny.value + 1) % 3, (ny.value + 2) % 3
            n = (x1[nP] - x0[nP] + 1) * (x1[nPP] - x0[nPP] + 1)
            k = [np.zeros(n, np.float32) for _ in range(4)]
            self._f.abc_coeff(self._h, a, *[v.ctypes.data_as(_fp) for v in k])
            out.append(dict(ny=ny.value, type=ty.value, positive=pos.value, x0=tuple(x0), x1=tuple(x1),
                            K1P=k[0], K1PP=k[1], K2P=k[2], K2PP=k[3]))
        return out

    def mur_planes(self):
        out = []
        for m in range(self._f.mur_count(self._h)):
            ny, top = C.c_int(), C.c_int()
            line, shift, st = C.c_uint(), C.c_uint(), C.c_uint()
            nl = (C.c_uint * 2)()
            self._f.mur_info(self._h, m, C.byref(ny), C.byref(top), C.byref(line), C.byref(shift), nl, C.byref(st))
            cnt = nl[0] * nl[1]
            cP = np.ctypeslib.as_array(self._f.mur_coeff(self._h, m, 0), shape=(cnt,)).reshape(nl[0], nl[1])
            cPP = np.ctypeslib.as_array(self._f.mur_coeff(self._h, m, 1), shape=(cnt,)).reshape(nl[0], nl[1])
            out.append(dict(ny=ny.value, top=top.value, line=line.value, shift=shift.value,
                            n=(nl[0], nl[1]), start_ts=st.value, coeff_nyP=cP, coeff_nyPP=cPP))
        return out

    def lorentz(self):
        out = []
        for o in range(self._f.lorentz_order(self._h)):
            cnt = self._f.lorentz_count(self._h, o)
            flags = self._f.lorentz_flags(self._h, o)
            pos = np.zeros((3, cnt), dtype=np.uint32)
            co = {}
            for n in range(3):
                if cnt:
                    pos[n] = np.ctypeslib.as_array(self._f.lorentz_pos(self._h, o, n), shape=(cnt,))
            for w, name in enumerate(("v_int", "v_ext", "v_lor", "i_int", "i_ext", "i_lor")):
                ptr = self._f.lorentz_coeff(self._h, o, w, 0)
                if not ptr or cnt == 0:
                    co[name] = None
                    continue
                a = np.zeros((3, cnt), dtype=np.float32)
                for n in range(3):
                    a[n] = np.ctypeslib.as_array(self._f.lorentz_coeff(self._h, o, w, n), shape=(cnt,))
                co[name] = a
            out.append(dict(count=cnt, flags=flags, pos=pos, **co))
        return out

    # ---- engine
    def iterate(self, n):
        self._f.iterate(self._h, n)

    @property
    def num_ts(self):
        return self._f.num_ts(self._h)

    @property
    def volt(self):
        return self._field_view(self._f.volt(self._h))

    @property
    def curr(self):
        return self._field_view(self._f.curr(self._h))

    def reset_fields(self):
        self._f.reset_fields(self._h)

    # ---- readout
    def voltage_integral(self, start, stop):
        return self._f.voltage_integral(self._h, _u3(*start), _u3(*stop))

    def current_integral(self, start, stop, norm_dir, start_inside=(1, 1, 1), stop_inside=(1, 1, 1)):
        return self._f.current_integral(self._h, _u3(*start), _u3(*stop), norm_dir,
                                          _i3(*start_inside), _i3(*stop_inside))

    def raw_field(self, is_H, pos):
        out = _d3()
        self._f.raw_field(self._h, int(is_H), _u3(*pos), out)
        return np.array(out[:])

    def energy(self):
        return self._f.energy(self._h)

    def dump_field(self, is_H, interp, start, stop):
        n = [stop[i] - start[i] + 1 for i in range(3)]
        out = np.zeros((3, n[2], n[1], n[0]), dtype=np.float32)
        self._f.dump_field(self._h, int(is_H), interp, _u3(*start), _u3(*stop), out.ctypes.data_as(_fp))
        return out

    def mode_match(self, is_H, ny, start, stop, dist0, dist1):
        """ProcessModeMatch::CalcMultipleIntegrals: (value, value^2/purity)"""
        d0 = np.ascontiguousarray(dist0, np.float64)
        d1 = np.ascontiguousarray(dist1, np.float64)
        out = np.zeros(2, np.float64)
        dp = C.POINTER(C.c_double)
        self._f.mode_match(self._h, int(is_H), int(ny), _u3(*start), _u3(*stop), d0.ctypes.data_as(dp), d1.ctypes.data_as(dp),
                             out.ctypes.data_as(dp))
        return float(out[0]), float(out[1])

    @staticmethod
    def fd_weight(freq, T, dT, interval):
        """exp_jwt_2_dt of processfields_fd.cpp:84-86 as complex64"""
        w = np.zeros(2, np.float32)
        lib().orc_fd_weight(float(freq), float(T), float(dT), int(interval), w.ctypes.data_as(_fp))
        return np.complex64(complex(w[0], w[1]))

    @staticmethod
    def fd_accumulate(acc, td, weight):
        """acc (complex64, any shape) += td (float32, same shape) * weight, processfields_fd.cpp:88-100"""
        assert acc.dtype == np.complex64 and td.dtype == np.float32 and acc.shape == td.shape and acc.flags.c_contiguous
        w = np.array([np.float32(weight.real), np.float32(weight.imag)], np.float32)
        tdc = np.ascontiguousarray(td)
        lib().orc_fd_accumulate(acc.view(np.float32).ctypes.data_as(_fp), tdc.ctypes.data_as(_fp), tdc.size, w.ctypes.data_as(_fp))

    def edge_length(self, n, pos, dual=False):
        return self._f.edge_length(self._h, n, _u3(*pos), int(dual))

    def disc_line(self, n, pos, dual=False):
        return self._f.disc_line(self._h, n, pos, int(dual))


class OracleSSE:
    """sse-compressed + x-slab multithreaded engine restatement (fdtd_oracle_sse.c) running on
    the operator of an OracleSim: the CPU baseline and the second leg of the bit-equality rule."""

    def __init__(self, sim: OracleSim, threads=1):
        self.sim = sim
        self._h = lib().orc_sse_create(sim._h, threads)
        if not self._h:
            raise RuntimeError("orc_sse_create failed")

    def close(self):
        if self._h:
            lib().orc_sse_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def iterate(self, n):
        lib().orc_sse_iterate(self._h, n)

    @property
    def unique(self):
        return lib().orc_sse_unique(self._h)

    @property
    def num_ts(self):
        return lib().orc_sse_num_ts(self._h)

    def fields(self):
        N = self.sim.N
        v = np.zeros((3,) + tuple(N), dtype=np.float32)
        c = np.zeros((3,) + tuple(N), dtype=np.float32)
        lib().orc_sse_get_fields(self._h, v.ctypes.data_as(_fp), c.ctypes.data_as(_fp))
        return v, c
