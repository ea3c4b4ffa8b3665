"""SyntheticOperator: Python face of the host-side operator builder in libopenems_b200.so
(csrc/host/synthetic_operator.cpp).  It plays the role of Operator::CalcECOperator +
Operator_CUDA's compression for box geometries on (non-)uniform Cartesian meshes and emits the
compressed device format directly, so 1024^3 meshes build in seconds within a few GB of host
memory.  Host code only; the time stepping is always on the GPU."""
import ctypes as C

import numpy as np

from ._lib import load_library, CoeffEntry
from .engine import Operator_CUDA, EngineError

_d3 = C.c_double * 3
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint)

BC_PEC, BC_PMC, BC_MUR, BC_PML = 0, 1, 2, 3
EXC_E_SOFT, EXC_E_HARD, EXC_H_SOFT, EXC_H_HARD = 0, 1, 2, 3


class SyntheticOperator:
    def __init__(self, x, y, z, grid_delta=1.0):
        self._L = load_library()
        self.x, self.y, self.z = (np.ascontiguousarray(a, np.float64) for a in (x, y, z))
        self.N = (len(self.x), len(self.y), len(self.z))
        self.grid_delta = float(grid_delta)
        self._h = self._L.oems_synth_create(*self.N, self.x.ctypes.data_as(_dp), self.y.ctypes.data_as(_dp),
                                            self.z.ctypes.data_as(_dp), grid_delta)
        if not self._h:
            raise EngineError("SyntheticOperator: need at least 3 lines per direction")
        self._built = False

    def close(self):
        if getattr(self, "_h", None):
            self._L.oems_synth_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bc(self, bc, pml_size=(8,) * 6):
        self._L.oems_synth_set_bc(self._h, (C.c_int * 6)(*bc), (C.c_uint * 6)(*pml_size))

    def set_background(self, epsR=1.0, mueR=1.0, kappa=0.0, sigma=0.0):
        self._L.oems_synth_set_background(self._h, epsR, mueR, kappa, sigma)

    def set_timestep(self, forced_dT=0.0, factor=1.0):
        self._L.oems_synth_set_timestep(self._h, forced_dT, factor)

    def add_material(self, start, stop, epsR=1.0, mueR=1.0, kappa=0.0, sigma=0.0, prio=0):
        return self._L.oems_synth_add_material(self._h, prio, _d3(*start), _d3(*stop), epsR, mueR, kappa, sigma)

    def add_metal(self, start, stop, prio=10):
        return self._L.oems_synth_add_metal(self._h, prio, _d3(*start), _d3(*stop))

    def add_lorentz(self, start, stop, epsR=1.0, mueR=1.0, kappa=0.0, sigma=0.0, prio=0,
                    eps_fp=(), eps_tau=(), eps_flor=(), mue_fp=(), mue_tau=(), mue_flor=()):
        order = max(len(eps_fp), len(mue_fp))

        def arr(v):
            a = np.zeros(max(order, 1), np.float64)
            a[:len(v)] = v
            return a
        arrs = [arr(v) for v in (eps_fp, eps_tau, eps_flor, mue_fp, mue_tau, mue_flor)]
        return self._L.oems_synth_add_lorentz(self._h, prio, _d3(*start), _d3(*stop), epsR, mueR, kappa, sigma,
                                              order, *[a.ctypes.data_as(_dp) for a in arrs])

    def add_excitation(self, start, stop, exc_type, vec, delay=0.0, prio=0):
        return self._L.oems_synth_add_excitation(self._h, prio, _d3(*start), _d3(*stop), exc_type, _d3(*vec), delay)

    def set_excite_gauss(self, f0, fc):
        self._L.oems_synth_set_excite_gauss(self._h, f0, fc)

    def set_excite_sinus(self, f0):
        self._L.oems_synth_set_excite_sinus(self._h, f0)

    def set_slab(self, z_begin, z_end):
        """one process per GPU: build only the planes of the owned range [z_begin, z_end) (+ ghost planes)"""
        if self._L.oems_synth_set_slab(self._h, int(z_begin), int(z_end)):
            raise EngineError((self._L.oems_synth_last_error(self._h) or b"").decode())

    def local_timestep(self):
        """Operator::CalcTimestep over this rank's planes; MIN-reduce over the ranks, then set_timestep(dT)"""
        d = C.c_double()
        if self._L.oems_synth_local_timestep(self._h, C.byref(d)):
            raise EngineError("oems_synth_local_timestep failed")
        return d.value

    def build(self, max_ts=10 ** 9):
        if self._L.oems_synth_build(self._h, max_ts):
            raise EngineError((self._L.oems_synth_last_error(self._h) or b"").decode())
        self._built = True

    # ---- results
    @property
    def dT(self): return self._L.oems_synth_dT(self._h)
    @property
    def nyquist(self): return self._L.oems_synth_nyquist(self._h)
    @property
    def n_unique(self): return self._L.oems_synth_n_unique(self._h)
    @property
    def index_bytes(self): return self._L.oems_synth_index_bytes(self._h)
    @property
    def unique_planes(self): return self._L.oems_synth_unique_planes(self._h)

    def table(self):
        """float32 view [U][32] of the coefficient tuples (oems_coeff_entry)"""
        n = self.n_unique
        ptr = C.cast(self._L.oems_synth_table(self._h), _fp)
        return np.ctypeslib.as_array(ptr, shape=(n, 32))

    def index(self):
        dt = np.uint16 if self.index_bytes == 2 else np.uint32
        cnt = self.N[0] * self.N[1] * self.N[2]
        ptr = C.cast(self._L.oems_synth_index(self._h), C.POINTER(C.c_uint16 if dt == np.uint16 else C.c_uint32))
        return np.ctypeslib.as_array(ptr, shape=(cnt,)).reshape(self.N[2], self.N[1], self.N[0])

    def planes(self):
        """the index as the engine receives it: (unique xy planes [P][ny][nx], plane id per z [nz])"""
        dt = np.uint16 if self.index_bytes == 2 else np.uint32
        P = self.unique_planes
        ptr = C.cast(self._L.oems_synth_plane_data(self._h), C.POINTER(C.c_uint16 if dt == np.uint16 else C.c_uint32))
        up = np.ctypeslib.as_array(ptr, shape=(P * self.N[1] * self.N[0],)).reshape(P, self.N[1], self.N[0])
        ids = np.ctypeslib.as_array(self._L.oems_synth_plane_of_z(self._h), shape=(self.N[2],))
        return up, ids

    def dense(self, which):
        """expands table[index] to an ArrayNIJK coefficient array (tests only; O(N) memory)"""
        col = {"vv": 0, "vi": 3, "ii": 6, "iv": 9, "pml": 12, "pml_vv": 13, "pml_vvfn": 16, "pml_vvfo": 19,
               "pml_ii": 22, "pml_iifn": 25, "pml_iifo": 28}[which]
        t, idx = self.table(), self.index()
        if which == "pml":
            return t[idx, col].transpose(2, 1, 0)
        return np.stack([t[idx, col + n].transpose(2, 1, 0) for n in range(3)])

    def signal(self):
        n = self._L.oems_synth_signal_length(self._h)
        return (np.ctypeslib.as_array(self._L.oems_synth_signal(self._h, 0), shape=(n,)).copy(),
                np.ctypeslib.as_array(self._L.oems_synth_signal(self._h, 1), shape=(n,)).copy())

    def excitation(self, is_curr):
        n = self._L.oems_synth_exc_count(self._h, int(is_curr))
        idx = np.zeros((3, n), np.uint32)
        d = np.zeros(n, np.uint32)
        amp = np.zeros(n, np.float32)
        delay = np.zeros(n, np.uint32)
        if n:
            self._L.oems_synth_exc_get(self._h, int(is_curr), idx.ctypes.data_as(_up), d.ctypes.data_as(_up),
                                       amp.ctypes.data_as(_fp), delay.ctypes.data_as(_up))
        return idx, d, amp, delay

    def mur_planes(self):
        out = []
        for m in range(self._L.oems_synth_mur_count(self._h)):
            ny = C.c_int()
            line, shift, st = C.c_uint(), C.c_uint(), C.c_uint()
            nl = (C.c_uint * 2)()
            p0 = self._L.oems_synth_mur_coeff(self._h, m, 0, C.byref(ny), C.byref(line), C.byref(shift), nl, C.byref(st))
            p1 = self._L.oems_synth_mur_coeff(self._h, m, 1, None, None, None, (C.c_uint * 2)(), None)
            cnt = nl[0] * nl[1]
            out.append(dict(ny=ny.value, line=line.value, shift=shift.value, n=(nl[0], nl[1]), start_ts=st.value,
                            coeff_nyP=np.ctypeslib.as_array(p0, shape=(cnt,)).reshape(nl[0], nl[1]).copy(),
                            coeff_nyPP=np.ctypeslib.as_array(p1, shape=(cnt,)).reshape(nl[0], nl[1]).copy()))
        return out

    def upml_boxes(self):
        out = []
        for b in range(self._L.oems_synth_upml_count(self._h)):
            st, nl = (C.c_uint * 3)(), (C.c_uint * 3)()
            self._L.oems_synth_upml_box(self._h, b, st, nl)
            out.append(dict(start=tuple(st), n=tuple(nl)))
        return out

    def lorentz_counts(self):
        return [self._L.oems_synth_lorentz_count(self._h, o) for o in range(self._L.oems_synth_lorentz_order(self._h))]

    def pin(self):
        """page-lock the operator index (once, after build): engine creation then copies it host ->
        device at the PCIe rate; returns False when there is no CUDA device to register with"""
        if not self._built:
            self.build()
        return self._L.oems_synth_pin(self._h) == 0

    # ---- hand-over to the engine: Operator::CreateEngine
    def operator(self):
        if not self._built:
            self.build()
        op = Operator_CUDA(self.N)
        op._synth = self
        op.SetTimestep(self.dT)
        op.SetMesh(self.x, self.y, self.z, self.grid_delta)
        return op

    def CreateEngine(self, device=-1, slab=None):
        return self.operator().CreateEngine(device=device, slab=slab)
