"""ctypes binding of oracle/_ref/libopenems_ref.so.  TEST INFRASTRUCTURE ONLY.

libopenems_ref.so is the reference's own code: the unmodified translation units of /root/reference
(engine, sse / sse-compressed / multithreaded engines, operator, all extensions, Processing classes)
compiled by oracle/Makefile.ref against the shim headers in oracle/ref_shim/ and driven by
oracle/ref_driver.cpp.  RefSim has the interface of pyoracle.OracleSim (same methods, same array
conventions), so one test case can be pushed through the CPU restatement, through the reference and
through the CUDA engine.  /root/reference only exists in the build container: on the GPU box the
prebuilt library (git-ignored, shipped by gpurun) is used as it is.

Only tests/, __graft_entry__ and bench.py's reference/cpu_baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import pyoracle
from .pyoracle import OracleSim, _Namespace, _u3, _d3, _dp, _fp, _up

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libopenems_ref.so")
_LIB_CUDA = os.path.join(_HERE, "_ref", "libopenems_ref_cuda.so")
REFERENCE_TREE = os.environ.get("OPENEMS_REFERENCE", "/root/reference")

ENGINE_BASIC, ENGINE_SSE, ENGINE_SSE_COMPRESSED, ENGINE_MULTITHREADED, ENGINE_CUDA = 0, 1, 2, 3, 4


def have_reference_tree():
    return os.path.exists(os.path.join(REFERENCE_TREE, "FDTD", "engine.cpp"))


def available():
    return os.path.exists(_LIB) or have_reference_tree()


def build(force=False):
    """compile oracle/_ref/libopenems_ref.so from the reference sources where they lie (only possible
    where /root/reference exists); otherwise return the prebuilt library"""
    if have_reference_tree():
        if force:
            subprocess.check_call(["make", "-s", "-f", "Makefile.ref", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
        subprocess.check_call(["make", "-s", "-f", "Makefile.ref", "-C", _HERE, "-j", str(os.cpu_count() or 4),
                               "REF=" + REFERENCE_TREE], stdout=subprocess.DEVNULL)
    if not os.path.exists(_LIB):
        raise FileNotFoundError("oracle/_ref/libopenems_ref.so missing and no reference tree at %s" % REFERENCE_TREE)
    return _LIB


_lib = None
_lib_cuda = None


def cuda_available():
    """the harness variant that also contains integration/*.cpp (Operator_CUDA, Engine_CUDA, ...) linked with
    libopenems_b200.so"""
    if have_reference_tree():
        build()
    return os.path.exists(_LIB_CUDA)


def lib_cuda():
    global _lib_cuda
    if _lib_cuda is None:
        build()
        if not os.path.exists(_LIB_CUDA):
            raise FileNotFoundError("oracle/_ref/libopenems_ref_cuda.so missing")
        _lib_cuda = _bind(C.CDLL(_LIB_CUDA))
        _lib_cuda.ref_set_fast_processing.restype = None
        _lib_cuda.ref_set_fast_processing.argtypes = [C.c_void_p, C.c_int]
    return _lib_cuda


def lib():
    global _lib
    if _lib is None:
        _lib = _bind(C.CDLL(build()))
    return _lib


def _bind(L):
    pyoracle.lib()   # fills pyoracle.SIGNATURES
    vp = C.c_void_p
    for name, (res, args) in pyoracle.SIGNATURES.items():
        rname = "ref_" + name[4:]
        if name.startswith("orc_sse_") or not hasattr(L, rname):
            continue
        f = getattr(L, rname)
        f.restype, f.argtypes = res, args
    extra = {
        "ref_set_engine": (None, [vp, C.c_int, C.c_int]),
        "ref_add_debye": (C.c_int, [vp, C.c_int, _d3, _d3] + [C.c_double] * 4 + [C.c_int, _dp, _dp]),
        "ref_add_conducting_sheet": (C.c_int, [vp, C.c_int, _d3, _d3, C.c_double, C.c_double]),
        "ref_add_lumped_rlc": (C.c_int, [vp, _d3, _d3, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]),
        "ref_set_excite_custom": (None, [vp, C.c_char_p, C.c_double, C.c_double]),
        "ref_set_field": (None, [vp, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_float]),
        "ref_lorentz_ext_count": (C.c_int, [vp]),
        "ref_lorentz_select": (None, [vp, C.c_int]),
        "ref_rlc_count": (C.c_uint, [vp]),
        "ref_rlc_get": (None, [vp, C.POINTER(C.c_int), _up, _fp]),
        "ref_add_probe": (C.c_int, [vp, C.c_int, C.c_char_p, _d3, _d3, C.c_double, C.c_int]),
        "ref_add_dump": (C.c_int, [vp, C.c_char_p, _d3, _d3, C.c_int, C.c_int, C.c_int, C.c_uint]),
        "ref_add_fd_dump": (C.c_int, [vp, C.c_char_p, _d3, _d3, C.c_int, C.c_int, C.c_uint, _dp]),
        "ref_add_mode_match": (C.c_int, [vp, C.c_char_p, _d3, _d3, C.c_int, C.c_char_p, C.c_char_p, C.c_int]),
        "ref_has_cuda": (C.c_int, []),
        "ref_run": (None, [vp, C.c_uint]),
        "ref_recorded_count": (C.c_int, []),
        "ref_recorded_key": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
        "ref_recorded_size": (C.c_long, [C.c_char_p]),
        "ref_recorded_dims": (C.c_int, [C.c_char_p, C.POINTER(C.c_ulong), C.c_int]),
        "ref_recorded_get": (C.c_int, [C.c_char_p, _dp]),
        "ref_recorded_clear": (None, []),
        "ref_sse_unique": (C.c_uint, [vp]),
        "ref_version": (C.c_char_p, []),
    }
    for name, (res, args) in extra.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


class RefSim(OracleSim):
    """The reference's own operator + engine classes behind the OracleSim interface.

    engine: ENGINE_BASIC (Engine, FDTD/engine.cpp), ENGINE_SSE, ENGINE_SSE_COMPRESSED,
    ENGINE_MULTITHREADED (Engine_Multithread, the reference's default)."""

    def __init__(self, x, y, z, grid_delta=1.0, engine=ENGINE_BASIC, threads=1, fast_processing=True, cuda_lib=False):
        self.engine = engine
        self.cuda_lib = bool(cuda_lib) or engine == ENGINE_CUDA   # which of the two harness libraries serves this sim
        super().__init__(x, y, z, grid_delta)
        self._f.set_engine(self._h, engine, threads)
        if engine == ENGINE_CUDA:
            self._f.set_fast_processing(self._h, int(fast_processing))

    def _functions(self):
        # ENGINE_CUDA: the reference's operator and Processing classes around integration/Engine_CUDA + libopenems_b200.so
        return _Namespace(lib_cuda() if self.cuda_lib else lib(), "ref_")

    def add_mode_match(self, name, start, stop, field_type, func_P, func_PP, ny):
        return self._f.add_mode_match(self._h, name.encode(), _d3(*start), _d3(*stop), field_type, func_P.encode(), func_PP.encode(), ny)

    # fields are snapshots here (the reference engines keep their own layouts)
    def set_field(self, is_curr, n, x, y, z, value):
        self._f.set_field(self._h, int(is_curr), n, x, y, z, value)

    def add_debye(self, start, stop, epsR=1.0, mueR=1.0, kappa=0.0, sigma=0.0, prio=0, eps_delta=(), eps_tau=()):
        d = np.ascontiguousarray(eps_delta, np.float64)
        t = np.ascontiguousarray(eps_tau, np.float64)
        return self._f.add_debye(self._h, prio, _d3(*start), _d3(*stop), epsR, mueR, kappa, sigma, len(d),
                                 d.ctypes.data_as(_dp), t.ctypes.data_as(_dp))

    def add_conducting_sheet(self, start, stop, conductivity, thickness, prio=10):
        return self._f.add_conducting_sheet(self._h, prio, _d3(*start), _d3(*stop), conductivity, thickness)

    def add_lumped_rlc(self, start, stop, direction, R=float("nan"), Cap=float("nan"), L=float("nan"), series=False, caps=True):
        return self._f.add_lumped_rlc(self._h, _d3(*start), _d3(*stop), direction, R, Cap, L, int(series), int(caps))

    def set_excite_custom(self, func, f0, fmax):
        self._f.set_excite_custom(self._h, func.encode(), f0, fmax)

    def lorentz_extensions(self):
        """tables of every dispersive extension (Lorentz/Drude/Debye first, conducting sheet second)"""
        out = []
        for e in range(self._f.lorentz_ext_count(self._h)):
            self._f.lorentz_select(self._h, e)
            out.append(self.lorentz())
        self._f.lorentz_select(self._h, 0)
        return out

    def rlc_tables(self):
        """lumped RLC tables as built by Operator_Ext_LumpedRLC::BuildExtension"""
        n = self._f.rlc_count(self._h)
        d = np.zeros(n, np.int32)
        pos = np.zeros((3, n), np.uint32)
        co = np.zeros((9, n), np.float32)
        if n:
            self._f.rlc_get(self._h, d.ctypes.data_as(C.POINTER(C.c_int)), pos.ctypes.data_as(_up), co.ctypes.data_as(_fp))
        names = ("ilv", "i2v", "vvd", "vv2", "vj1", "vj2", "ib0", "b1", "b2")
        return d, pos, {k: co[i] for i, k in enumerate(names)}

    @property
    def sse_unique(self):
        return self._f.sse_unique(self._h)

    # ---- the reference's Processing classes
    def add_probe(self, kind, name, start, stop, weight=1.0, norm_dir=-1):
        """kind 0 voltage 1 current 2 E-field 3 H-field probe; coordinates in drawing units; writes the ASCII
        series to the file `name` (cwd) exactly like openEMS"""
        return self._f.add_probe(self._h, kind, name.encode(), _d3(*start), _d3(*stop), weight, norm_dir)

    def add_dump(self, name, start, stop, dump_type=0, file_type=1, interp=0, interval=0):
        return self._f.add_dump(self._h, name.encode(), _d3(*start), _d3(*stop), dump_type, file_type, interp, interval)

    def add_fd_dump(self, name, start, stop, freqs, dump_type=0, interp=0):
        f = np.ascontiguousarray(freqs, np.float64)
        return self._f.add_fd_dump(self._h, name.encode(), _d3(*start), _d3(*stop), dump_type, interp, len(f), f.ctypes.data_as(_dp))

    def run(self, nr_ts):
        """openEMS::RunFDTD's loop: bursts of IterateTS between Processing::Process calls"""
        self._f.run(self._h, nr_ts)


def recorded(cuda=False):
    """datasets the reference's HDF5/VTK writers were asked to write: {key: ndarray}"""
    L = lib_cuda() if cuda else lib()
    out = {}
    buf = C.create_string_buffer(1024)
    for i in range(L.ref_recorded_count()):
        L.ref_recorded_key(i, buf, 1024)
        key = buf.value
        dims = (C.c_ulong * 8)()
        nd = L.ref_recorded_dims(key, dims, 8)
        a = np.zeros(L.ref_recorded_size(key), np.float64)
        L.ref_recorded_get(key, a.ctypes.data_as(_dp))
        out[key.decode()] = a.reshape([dims[d] for d in range(nd)])
    return out


def recorded_clear(cuda=False):
    (lib_cuda() if cuda else lib()).ref_recorded_clear()


def read_probe_file(path):
    """time / value columns of a probe file written by ProcessIntegral (Common/processintegral.cpp:134-171)"""
    return np.loadtxt(path, comments="%", ndmin=2)
