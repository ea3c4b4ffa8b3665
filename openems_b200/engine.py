"""Host-side mirror of the reference's Operator / Engine / Engine_Interface API for the hot
path, bound to libopenems_b200.so through ctypes.

Reference interfaces mirrored (file:line relative to the openEMS tree):
  Operator      FDTD/operator.h:56-126,215-218   SetVV/GetVV..., GetNumberOfLines, CreateEngine
  Engine        FDTD/engine.h:41-125             IterateTS, GetNumberOfTimesteps, Get/Set Volt/Curr
  Engine_Interface_Base  Common/engine_interface_base.h:52-82  CalcVoltageIntegral, GetEField, ...
Errors: the reference prints to cerr and returns false / exits; here every failing C-ABI call
raises EngineError carrying oems_cuda_last_error().
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CoeffEntry, Stats, load_library

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)
_ip = C.POINTER(C.c_int)
_u3 = C.c_uint * 3
_i3 = C.c_int * 3


class EngineError(RuntimeError):
    pass


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Operator_CUDA:
    """Holds what Operator + its Operator_Extensions hold after CalcECOperator() on the host
    and uploads it in CreateEngine() (the single H2D crossing)."""

    def __init__(self, numLines):
        self.numLines = tuple(int(n) for n in numLines)
        shape = (3,) + self.numLines
        self.vv = None
        self.vi = None
        self.ii = None
        self.iv = None
        self._shape = shape
        self._compressed = None
        self._synth = None
        self.dT = 0.0
        self.signal = None
        self.exc = {0: None, 1: None}
        self.upml = []
        self.mur = []
        self.lorentz = []
        self.rlc = []
        self.sheets = []
        self.tfsf = None
        self.steadystate = None
        self.mesh = None  # (x, y, z, gridDelta) for field probes / dumps

    # ---- Operator::GetNumberOfLines / SetVV.. / GetVV.. (FDTD/operator.h:215-218)
    def GetNumberOfLines(self, n, full=True):
        return self.numLines[n]

    def _dense(self):
        if self.vv is None:
            self.vv = np.zeros(self._shape, np.float32)
            self.vi = np.zeros(self._shape, np.float32)
            self.ii = np.zeros(self._shape, np.float32)
            self.iv = np.zeros(self._shape, np.float32)

    def SetVV(self, n, x, y, z, v): self._dense(); self.vv[n, x, y, z] = v
    def SetVI(self, n, x, y, z, v): self._dense(); self.vi[n, x, y, z] = v
    def SetII(self, n, x, y, z, v): self._dense(); self.ii[n, x, y, z] = v
    def SetIV(self, n, x, y, z, v): self._dense(); self.iv[n, x, y, z] = v
    def GetVV(self, n, x, y, z): return self.vv[n, x, y, z]
    def GetVI(self, n, x, y, z): return self.vi[n, x, y, z]
    def GetII(self, n, x, y, z): return self.ii[n, x, y, z]
    def GetIV(self, n, x, y, z): return self.iv[n, x, y, z]

    def SetOperatorArrays(self, vv, vi, ii, iv):
        """dense ArrayNIJK coefficient arrays [3][Nx][Ny][Nz]"""
        for a in (vv, vi, ii, iv):
            if tuple(a.shape) != self._shape:
                raise EngineError("operator array shape %s != %s" % (a.shape, self._shape))
        self.vv, self.vi, self.ii, self.iv = _f32(vv), _f32(vi), _f32(ii), _f32(iv)

    def SetCompressedOperator(self, table, index):
        """table: array of CoeffEntry (or float32 [U][32]); index [Nz][Ny][Nx] uint16/uint32"""
        self._compressed = (table, index)

    def SetTimestep(self, dT): self.dT = float(dT)
    def GetTimestep(self): return self.dT

    def SetMesh(self, x, y, z, gridDelta=1.0):
        self.mesh = (np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(z, np.float64), float(gridDelta))

    # ---- extension data (Operator_Ext_*)
    def SetExcitationSignal(self, sig_volt, sig_curr, period_ts=0):
        self.signal = (_f32(sig_volt), _f32(sig_curr), int(period_ts))

    def SetExcitation(self, is_curr, index3, direction, amp, delay):
        self.exc[int(bool(is_curr))] = (_u32(index3), _u32(direction), _f32(amp), _u32(delay))

    def AddUPML(self, start, numLines, vv, vvfn, vvfo, ii, iifn, iifo):
        self.upml.append((tuple(start), tuple(numLines), [None if a is None else _f32(a) for a in (vv, vvfn, vvfo, ii, iifn, iifo)]))

    def AddMur(self, ny, lineNr, lineNrShift, coeff_nyP, coeff_nyPP, start_TS=0):
        self.mur.append((int(ny), int(lineNr), int(lineNrShift), _f32(coeff_nyP), _f32(coeff_nyPP), int(start_TS)))

    def AddLorentzOrder(self, pos3, v_int=None, v_ext=None, v_lor=None, i_int=None, i_ext=None, i_lor=None):
        self.lorentz.append((_u32(pos3), [None if a is None else _f32(a) for a in (v_int, v_ext, v_lor, i_int, i_ext, i_lor)]))

    def SetSteadyStateDetection(self, period_ts, pos3=None, direction=None):
        """Operator_Ext_SteadyState; without an explicit probe list the driver's default set is
        used (openems.cpp:1206-1234: centre node plus, per axis, the node with that coordinate
        set to 0 -- the `pos[n] *= 1/4` of the reference is integer arithmetic -- twice)"""
        if pos3 is None:
            c = [n // 2 for n in self.numLines]
            pts = [tuple(c)]
            for n in range(3):
                q = list(c)
                q[n] = 0
                pts += [tuple(q), tuple(q)]
            pos3 = np.array([[p[a] for p in pts for _ in range(3)] for a in range(3)], np.uint32)
            direction = np.array([d for _ in pts for d in range(3)], np.uint32)
        self.steadystate = (int(period_ts), _u32(pos3), _u32(direction))

    def SetTFSF(self, start, stop, active, faces):
        """Operator_Ext_TFSF tables: active[n][l]; faces[(which, n, l, c)] = (delay uint32, delay_delta f32, amp f32)
        with which 0 = voltage, 1 = current"""
        self.tfsf = (_u32(start), _u32(stop), np.ascontiguousarray(np.array(active, np.int32).reshape(6)),
                     {k: (_u32(v[0]), _f32(v[1]), _f32(v[2])) for k, v in faces.items()})

    def AddAbsorbingSheet(self, ny, x0, x1, normal_positive, abc_type, K1P, K1PP, K2P=None, K2PP=None):
        """Operator_Ext_Absorbing_BC: sheet on mesh indices x0..x1 normal to ny, coefficient tables [nl0][nl1]"""
        self.sheets.append((int(ny), _u32(x0), _u32(x1), int(bool(normal_positive)), int(abc_type), _f32(K1P), _f32(K1PP),
                            None if K2P is None else _f32(K2P), None if K2PP is None else _f32(K2PP)))

    def AddLumpedRLC(self, direction, pos3, coeffs):
        names = ("ilv", "i2v", "vvd", "vv2", "vj1", "vj2", "ib0", "b1", "b2")
        self.rlc.append((np.ascontiguousarray(direction, np.int32), _u32(pos3), [_f32(coeffs[k]) for k in names]))

    # ---- Operator::CreateEngine (FDTD/operator.h:56)
    def CreateEngine(self, device=-1, slab=None, finalize=True):
        eng = Engine_CUDA(self, device=device, slab=slab)
        if finalize:
            eng.Init()
        return eng


class Engine_CUDA:
    """Engine (FDTD/engine.h:41-125) on one B200.  GetType() reports 'CUDA'."""

    def __init__(self, op: Operator_CUDA, device=-1, slab=None):
        self._L = load_library()
        self.Op = op
        self.numLines = op.numLines
        h = C.c_void_p()
        rc = self._L.oems_cuda_create(*op.numLines, device, C.byref(h))
        if rc:
            raise EngineError((self._L.oems_cuda_last_error(None) or b"").decode())
        self._h = h
        self._slab = slab
        if slab is not None:
            self._ck(self._L.oems_cuda_set_slab(self._h, slab[0], slab[1]))
        self._initialized = False
        self._probe_values = 0

    def _ck(self, rc):
        if rc:
            raise EngineError((self._L.oems_cuda_last_error(self._h) or b"unknown error").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.oems_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Engine::Init (engine.cpp:51-59): upload + extension schedule
    def Init(self):
        if self._initialized:
            return
        L, op, h = self._L, self.Op, self._h
        if op._synth is not None:
            self._ck(L.oems_synth_upload(op._synth._h, h))
        elif op._compressed is not None:
            table, index = op._compressed
            index = np.ascontiguousarray(index)
            if index.dtype not in (np.uint16, np.uint32):
                raise EngineError("index must be uint16 or uint32")
            if isinstance(table, np.ndarray):
                table = np.ascontiguousarray(table, dtype=np.float32)
                nu, tp = table.shape[0], table.ctypes.data_as(C.c_void_p)
            else:
                nu, tp = len(table), C.cast(table, C.c_void_p)
            self._keep = (table, index)
            self._ck(L.oems_cuda_set_operator_compressed(h, nu, tp, index.ctypes.data_as(C.c_void_p), index.dtype.itemsize))
        else:
            if op.vv is None:
                raise EngineError("Operator_CUDA holds no coefficients")
            self._ck(L.oems_cuda_set_operator_dense(h, _ptr(op.vv, _fp), _ptr(op.vi, _fp), _ptr(op.ii, _fp), _ptr(op.iv, _fp)))
        if op._synth is None:
            if op.signal is not None:
                sv, si, per = op.signal
                self._ck(L.oems_cuda_set_signal(h, _ptr(sv, _fp), _ptr(si, _fp), len(sv), per))
            for w in (0, 1):
                if op.exc[w] is not None:
                    idx, d, amp, delay = op.exc[w]
                    if len(d):
                        self._ck(L.oems_cuda_add_excitation(h, w, len(d), _ptr(idx, _up), _ptr(d, _up), _ptr(amp, _fp), _ptr(delay, _up)))
            for ny, line, shift, cP, cPP, st in op.mur:
                n2 = (C.c_uint * 2)(*cP.shape)
                self._ck(L.oems_cuda_add_mur(h, ny, line, shift, n2, _ptr(cP, _fp), _ptr(cPP, _fp), st))
            for start, nl, co in op.upml:
                self._ck(L.oems_cuda_add_upml(h, _u3(*start), _u3(*nl), *[_ptr(a, _fp) for a in co]))
            for pos3, co in op.lorentz:
                self._ck(L.oems_cuda_add_lorentz(h, pos3.shape[1], _ptr(pos3, _up), *[_ptr(a, _fp) for a in co]))
            for d, pos3, co in op.rlc:
                self._ck(L.oems_cuda_add_rlc(h, len(d), _ptr(d, _ip), _ptr(pos3, _up), *[_ptr(a, _fp) for a in co]))
            if op.tfsf is not None:
                st, sp, act, faces = op.tfsf
                arr = {}
                for which in (0, 1):
                    for kind, ctype in ((0, _up), (1, _fp), (2, _fp)):
                        a = (ctype * 12)()
                        for n in range(3):
                            for l in range(2):
                                for c in range(2):
                                    f = faces.get((which, n, l, c))
                                    if f is not None:
                                        a[(n * 2 + l) * 2 + c] = _ptr(f[kind], ctype)
                        arr[(which, kind)] = a
                self._ck(L.oems_cuda_set_tfsf(h, _ptr(st, _up), _ptr(sp, _up), act.ctypes.data_as(C.POINTER(C.c_int)),
                                              arr[(0, 0)], arr[(0, 1)], arr[(0, 2)], arr[(1, 0)], arr[(1, 1)], arr[(1, 2)]))
            for ny, x0, x1, pos, ty, k1p, k1pp, k2p, k2pp in op.sheets:
                self._ck(L.oems_cuda_add_absorbing_sheet(h, ny, _ptr(x0, _up), _ptr(x1, _up), pos, ty, _ptr(k1p, _fp), _ptr(k1pp, _fp),
                                                         None if k2p is None else _ptr(k2p, _fp), None if k2pp is None else _ptr(k2pp, _fp)))
            if op.steadystate is not None:
                per, pos3, d = op.steadystate
                self._ss_period = per
                self._ck(L.oems_cuda_add_steadystate(h, per, len(d), _ptr(pos3, _up), _ptr(d, _up)))
            self._ck(L.oems_cuda_finalize(h))
        self._initialized = True

    def GetType(self):
        return "CUDA"

    # ---- Engine::IterateTS / GetNumberOfTimesteps / NextInterval (engine.h:48-52)
    def IterateTS(self, iterTS):
        self._ck(self._L.oems_cuda_iterate(self._h, int(iterTS)))
        return True

    def IterateTimed(self, iterTS):
        """IterateTS + the burst's device time in ms (CUDA events on the engine's own stream)"""
        ms = C.c_double()
        self._ck(self._L.oems_cuda_iterate_timed(self._h, int(iterTS), C.byref(ms)))
        return ms.value

    def Synchronize(self):
        self._ck(self._L.oems_cuda_sync(self._h))

    def GetNumberOfTimesteps(self):
        ts = C.c_uint()
        self._ck(self._L.oems_cuda_num_ts(self._h, C.byref(ts)))
        return ts.value

    def NextInterval(self, curr_speed):
        return None  # thread auto-tuning hook of Engine_Multithread; nothing to tune here

    def Reset(self):
        self._ck(self._L.oems_cuda_reset(self._h))

    # ---- per-cell accessors (slow path), engine.h:55-101
    @staticmethod
    def _pos(args):
        if len(args) == 2:
            n, pos = args
            return int(n), int(pos[0]), int(pos[1]), int(pos[2])
        n, x, y, z = args
        return int(n), int(x), int(y), int(z)

    def GetVolt(self, *args):
        v = C.c_float()
        self._ck(self._L.oems_cuda_get_field(self._h, 0, *self._pos(args), C.byref(v)))
        return v.value

    def GetCurr(self, *args):
        v = C.c_float()
        self._ck(self._L.oems_cuda_get_field(self._h, 1, *self._pos(args), C.byref(v)))
        return v.value

    def SetVolt(self, *args):
        self._ck(self._L.oems_cuda_set_field(self._h, 0, *self._pos(args[:-1]), float(args[-1])))

    def SetCurr(self, *args):
        self._ck(self._L.oems_cuda_set_field(self._h, 1, *self._pos(args[:-1]), float(args[-1])))

    # ---- bulk access, ArrayNIJK order over the planes this engine holds
    def _held_planes(self):
        nz = self.numLines[2]
        if self._slab is None:
            return 0, nz
        zb, ze = self._slab
        return zb - (1 if zb > 0 else 0), ze + (1 if ze < nz else 0)

    def GetFields(self, is_curr):
        a, b = self._held_planes()
        out = np.zeros((3, self.numLines[0], self.numLines[1], b - a), np.float32)
        self._ck(self._L.oems_cuda_get_fields(self._h, int(is_curr), _ptr(out, _fp)))
        return out

    def SetFields(self, is_curr, arr):
        a, b = self._held_planes()
        arr = _f32(arr)
        if arr.shape != (3, self.numLines[0], self.numLines[1], b - a):
            raise EngineError("SetFields: wrong shape")
        self._ck(self._L.oems_cuda_set_fields(self._h, int(is_curr), _ptr(arr, _fp)))

    def GetUPMLFlux(self, box, is_curr, shape):
        out = np.zeros((3,) + tuple(shape), np.float32)
        self._ck(self._L.oems_cuda_get_upml_flux(self._h, box, int(is_curr), _ptr(out, _fp)))
        return out

    # ---- probes / energy / dumps
    def AddVoltageProbe(self, start, stop):
        i = C.c_int()
        self._ck(self._L.oems_cuda_add_probe_voltage(self._h, _u3(*start), _u3(*stop), C.byref(i)))
        return i.value

    def AddCurrentProbe(self, start, stop, normDir, start_inside=(1, 1, 1), stop_inside=(1, 1, 1)):
        i = C.c_int()
        self._ck(self._L.oems_cuda_add_probe_current(self._h, _u3(*start), _u3(*stop), normDir,
                                                     _i3(*[int(b) for b in start_inside]), _i3(*[int(b) for b in stop_inside]), C.byref(i)))
        return i.value

    def AddFieldProbe(self, is_H, pos):
        i = C.c_int()
        self._ck(self._L.oems_cuda_add_probe_field(self._h, int(is_H), _u3(*pos), C.byref(i)))
        return i.value

    def ReadProbes(self):
        """all probe values at the current timestep, flat, in probe-id order"""
        n = C.c_uint()
        out = np.zeros(4096, np.float64)
        self._ck(self._L.oems_cuda_read_probes(self._h, _ptr(out, _dp)))
        self._ck(self._L.oems_cuda_num_probe_values(self._h, C.byref(n)))
        return out[: n.value].copy()

    def RecordProbes(self, interval, max_samples):
        self._ck(self._L.oems_cuda_record_probes(self._h, int(interval), int(max_samples)))
        self._rec_cap = int(max_samples)

    def ReadProbeSeries(self):
        n = C.c_uint()
        nv = C.c_uint()
        self._ck(self._L.oems_cuda_num_probe_values(self._h, C.byref(nv)))
        cap = getattr(self, "_rec_cap", 0)
        out = np.zeros((max(cap, 1), max(nv.value, 1)), np.float64)
        ts = np.zeros(max(cap, 1), np.uint32)
        self._ck(self._L.oems_cuda_read_probe_series(self._h, _ptr(out, _dp), _ptr(ts, _up), cap, C.byref(n)))
        return ts[: n.value].copy(), out[: n.value].copy()

    def SteadyStateLastDiff(self):
        """Engine_Ext_SteadyState::GetLastDiff(): (last_max_diff, number of completed period checks)"""
        d, n = C.c_double(), C.c_uint()
        self._ck(self._L.oems_cuda_steadystate_check(self._h, C.byref(d), C.byref(n)))
        return d.value, n.value

    def CalcFastEnergy(self):
        e = C.c_double()
        self._ck(self._L.oems_cuda_energy(self._h, C.byref(e)))
        return e.value

    def AddDump(self, is_H, interp, px, py, pz, edge_len, dual_edge_len):
        px, py, pz = _u32(px), _u32(py), _u32(pz)
        el = [np.ascontiguousarray(a, np.float64) for a in edge_len]
        dl = [np.ascontiguousarray(a, np.float64) for a in dual_edge_len]
        elp = (_dp * 3)(*[a.ctypes.data_as(_dp) for a in el])
        dlp = (_dp * 3)(*[a.ctypes.data_as(_dp) for a in dl])
        i = C.c_int()
        self._ck(self._L.oems_cuda_add_dump(self._h, int(is_H), int(interp), len(px), len(py), len(pz),
                                            _ptr(px, _up), _ptr(py, _up), _ptr(pz, _up), elp, dlp, C.byref(i)))
        if not hasattr(self, "_dump_shapes"):
            self._dump_shapes = {}
        first, n = C.c_uint(), C.c_uint()
        self._ck(self._L.oems_cuda_dump_own_range(self._h, i.value, C.byref(first), C.byref(n)))
        self._dump_shapes[i.value] = (3, n.value, len(py), len(px))   # a z-slab engine dumps the z lines it owns
        if not hasattr(self, "_dump_own"):
            self._dump_own = {}
        self._dump_own[i.value] = (first.value, n.value)
        return i.value

    def DumpOwnRange(self, dump_id):
        """(first, n): the entries of the dump's z list this engine evaluates (all of them on a single GPU)"""
        return self._dump_own[dump_id]

    def ExchangeGhosts(self):
        """z-slab engines: complete the neighbours' ghost planes for an interpolating readout (enqueue only; collective)"""
        self._ck(self._L.oems_cuda_exchange_ghosts(self._h))

    def ReleaseGhosts(self):
        self._ck(self._L.oems_cuda_release_ghosts(self._h))

    def ReadDump(self, dump_id):
        out = np.zeros(self._dump_shapes[dump_id], np.float32)
        if out.size:
            self._ck(self._L.oems_cuda_read_dump(self._h, dump_id, _ptr(out, _fp)))
        else:   # a slab that owns none of the box still takes part in the ghost exchange
            self.ExchangeGhosts()
        return out

    def FillFields(self, seed=0):
        """deterministic pre-fill of E and H (function of the global cell index), see oems_cuda_fill_fields"""
        self._ck(self._L.oems_cuda_fill_fields(self._h, int(seed)))

    def FieldDigest(self):
        """(digest_E, digest_H) of the owned cells; slab digests add up (mod 2^64) to the single-GPU digest"""
        out = []
        for w in (0, 1):
            d = C.c_ulonglong()
            self._ck(self._L.oems_cuda_field_digest(self._h, w, C.byref(d)))
            out.append(d.value)
        return tuple(out)

    def ReadDumpAsync(self, dump_id):
        """ProcessFieldsTD::Process without stalling the time loop: evaluates the dump at the current timestep and starts
        its copy into page-locked host memory on a second stream; returns a ticket.  WaitDump(ticket) -> ndarray"""
        shape = self._dump_shapes[dump_id]
        if not hasattr(self, "_dump_pinned"):
            self._dump_pinned = {}
        if dump_id not in self._dump_pinned:
            p = C.c_void_p()
            if self._L.oems_cuda_host_alloc(max(1, int(np.prod(shape))) * 4, C.byref(p)):
                raise EngineError("oems_cuda_host_alloc failed")
            self._dump_pinned[dump_id] = p
        t = C.c_longlong()
        self._ck(self._L.oems_cuda_read_dump_async(self._h, dump_id, self._dump_pinned[dump_id], C.byref(t)))
        return (dump_id, t.value)

    def WaitDump(self, ticket):
        dump_id, t = ticket
        self._ck(self._L.oems_cuda_wait(self._h, t))
        shape = self._dump_shapes[dump_id]
        if not int(np.prod(shape)):
            return np.zeros(shape, np.float32)
        buf = (C.c_float * int(np.prod(shape))).from_address(self._dump_pinned[dump_id].value)
        return np.frombuffer(buf, np.float32).reshape(shape).copy()

    def AddFDDump(self, dump_id, n_freq):
        """ProcessFieldsFD::InitProcess: complex accumulators for n_freq frequencies on the device"""
        i = C.c_int()
        self._ck(self._L.oems_cuda_add_fd_dump(self._h, dump_id, int(n_freq), C.byref(i)))
        if not hasattr(self, "_fd_shapes"):
            self._fd_shapes = {}
        self._fd_shapes[i.value] = (int(n_freq),) + self._dump_shapes[dump_id]
        return i.value

    def AccumulateFD(self, fd_id, weights):
        """ProcessFieldsFD::Process for one sample; weights = complex64 exp_jwt_2_dt per frequency
        (processfields_fd.cpp:84-86)"""
        w = np.ascontiguousarray(weights, np.complex64)
        if w.shape != (self._fd_shapes[fd_id][0],):
            raise ValueError("one weight per frequency")
        self._ck(self._L.oems_cuda_fd_accumulate(self._h, fd_id, _ptr(w.view(np.float32), _fp)))

    def ReadFD(self, fd_id):
        """the accumulated spectra, complex64 [n_freq][3][nz][ny][nx], and the sample count"""
        out = np.zeros(self._fd_shapes[fd_id], np.complex64)
        n = C.c_uint()
        self._ck(self._L.oems_cuda_read_fd(self._h, fd_id, _ptr(out.view(np.float32), _fp) if out.size else None, C.byref(n)))
        return out, n.value

    def AddModeMatch(self, is_H, ny, start, stop, dist0, dist1, area, edge_len, dual_edge_len):
        """ProcessModeMatch::InitProcess with the host-evaluated, normalised mode template"""
        st, sp = _u32(start), _u32(stop)
        d0, d1, ar = (np.ascontiguousarray(a, np.float64) for a in (dist0, dist1, area))
        el = [np.ascontiguousarray(a, np.float64) for a in edge_len]
        dl = [np.ascontiguousarray(a, np.float64) for a in dual_edge_len]
        elp = (_dp * 3)(*[a.ctypes.data_as(_dp) for a in el])
        dlp = (_dp * 3)(*[a.ctypes.data_as(_dp) for a in dl])
        i = C.c_int()
        self._ck(self._L.oems_cuda_add_mode_match(self._h, int(is_H), int(ny), _ptr(st, _up), _ptr(sp, _up), _ptr(d0, _dp),
                                                  _ptr(d1, _dp), _ptr(ar, _dp), elp, dlp, C.byref(i)))
        return i.value

    def ReadModeMatch(self, mode_id):
        """ProcessModeMatch::CalcMultipleIntegrals: (value, value^2/purity)"""
        out = np.zeros(2, np.float64)
        self._ck(self._L.oems_cuda_read_mode_match(self._h, mode_id, _ptr(out, _dp)))
        return float(out[0]), float(out[1])

    def ReadModeMatchRaw(self, mode_id):
        """(value, value^2/purity, purity) over the planes this engine owns (z-slab partial result)"""
        out = np.zeros(3, np.float64)
        self._ck(self._L.oems_cuda_read_mode_match_raw(self._h, mode_id, _ptr(out, _dp)))
        return float(out[0]), float(out[1]), float(out[2])

    def SteadyStateRaw(self):
        """(info[2], energy[4], snap[2*period][count]) of this engine: see oems_cuda_steadystate_raw"""
        info = np.zeros(2, np.uint32)
        en = np.zeros(4, np.float64)
        cnt = C.c_uint()
        self._ck(self._L.oems_cuda_steadystate_raw(self._h, _ptr(info, _up), _ptr(en, _dp), None, 0, C.byref(cnt)))
        period = self._ss_period
        snap = np.zeros((2 * period, cnt.value), np.float64)
        if cnt.value:
            self._ck(self._L.oems_cuda_steadystate_raw(self._h, _ptr(info, _up), _ptr(en, _dp), _ptr(snap, _dp), snap.size, C.byref(cnt)))
        return info, en, snap

    def GetStats(self):
        s = Stats()
        self._ck(self._L.oems_cuda_get_stats(self._h, C.byref(s)))
        return dict(n_unique=s.n_unique, index_bytes=s.index_bytes, hbm_bytes=s.hbm_bytes,
                    kernels_launched=s.kernels_launched, kernels_per_step=s.kernels_per_step,
                    pml_cells=(s.pml_cells_hi << 32) | s.pml_cells_lo, uses_graph=bool(s.uses_graph))

    def SetTuning(self, block_rows=0, z_chunk=0, use_graph=-1):
        self._ck(self._L.oems_cuda_set_tuning(self._h, block_rows, z_chunk, use_graph))

    def SetOption(self, key, value):
        self._ck(self._L.oems_cuda_set_option(self._h, key.encode(), int(value)))

    def GetOption(self, key):
        v = C.c_longlong()
        self._ck(self._L.oems_cuda_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def TimeSchedule(self, n_ts):
        """average ms of every kernel of the per-timestep schedule (CUDA events on the engine stream)"""
        ms = np.zeros(64, np.float64)
        n = C.c_uint()
        self._ck(self._L.oems_cuda_time_schedule(self._h, int(n_ts), _ptr(ms, _dp), 64, C.byref(n)))
        return [((self._L.oems_cuda_schedule_label(self._h, i) or b"").decode(), float(ms[i])) for i in range(n.value)]

    # ---- multi-GPU
    def ExportIPC(self):
        buf = (C.c_ubyte * _lib.OEMS_IPC_BYTES)()
        self._ck(self._L.oems_cuda_export_ipc(self._h, buf))
        return bytes(buf)

    def OpenPeers(self, lower=None, upper=None):
        def conv(b):
            return None if b is None else (C.c_ubyte * _lib.OEMS_IPC_BYTES).from_buffer_copy(b)
        self._ck(self._L.oems_cuda_open_peers(self._h, conv(lower), conv(upper)))

    def LinkPeers(self, lower=None, upper=None):
        self._ck(self._L.oems_cuda_link_peers(self._h, lower._h if lower else None, upper._h if upper else None))


class Engine_Interface_CUDA:
    """Engine_Interface_FDTD (FDTD/engine_interface_fdtd.cpp) served from the device engine.
    Probes are registered on first use and then read in one batched device reduction."""

    def __init__(self, op: Operator_CUDA, eng: Engine_CUDA):
        self.Op, self.Eng = op, eng
        self._v = {}
        self._f = {}

    def GetNumberOfTimesteps(self):
        return self.Eng.GetNumberOfTimesteps()

    def GetTime(self, dualTime=False):
        return (self.Eng.GetNumberOfTimesteps() + (0.5 if dualTime else 0.0)) * self.Op.GetTimestep()

    def CalcVoltageIntegral(self, start, stop):
        key = (tuple(start), tuple(stop))
        if key not in self._v:
            self._v[key] = self.Eng.AddVoltageProbe(start, stop)
        return self._value(self._v[key])

    def _slots(self):
        # value slot of every probe id: voltage/current 1 value, field probes 3
        return None

    def _value(self, pid, n=1):
        vals = self.Eng.ReadProbes()
        # probe ids map to value offsets in registration order; recompute the offset table
        off = 0
        for kind_n, i in sorted(self._all_ids()):
            if i == pid:
                return vals[off] if kind_n == 1 else vals[off: off + 3]
            off += kind_n
        raise EngineError("unknown probe id")

    def _all_ids(self):
        ids = [(1, i) for i in self._v.values()] + [(3, i) for i in self._f.values()]
        return sorted(ids, key=lambda t: t[1])

    def _edge_length(self, n, pos, dual):
        x, y, z, gd = self.Op.mesh
        L = (x, y, z)[n]
        p, N = pos[n], len(L)
        if not dual:
            d = L[p + 1] - L[p] if p < N - 1 else L[p] - L[p - 1]
        else:
            def dl(q):
                return 0.5 * (L[q] + L[q + 1]) if q < N - 1 else L[q] + 0.5 * (L[q] - L[q - 1])
            d = dl(p) - dl(p - 1) if p > 0 else L[1] - L[0]
        return d * gd

    def _raw(self, is_H, pos):
        key = (int(is_H), tuple(pos))
        if key not in self._f:
            self._f[key] = self.Eng.AddFieldProbe(is_H, pos)
        raw = self._value(self._f[key], 3)
        out = np.zeros(3)
        for n in range(3):
            d = self._edge_length(n, pos, bool(is_H))
            out[n] = raw[n] / d if d else 0.0
        return out

    def GetEField(self, pos):
        return self._raw(0, pos)

    def GetHField(self, pos):
        return self._raw(1, pos)

    def CalcFastEnergy(self):
        return self.Eng.CalcFastEnergy()
