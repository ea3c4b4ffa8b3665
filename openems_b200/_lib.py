"""ctypes binding of libopenems_b200.so.  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.environ.get("OPENEMS_B200_LIB") or os.path.join(_HERE, "lib", "libopenems_b200.so")

OEMS_IPC_BYTES = 512


class LibraryNotBuilt(RuntimeError):
    pass


class CoeffEntry(C.Structure):
    """oems_coeff_entry (include/openems_b200.h)"""
    _fields_ = [("vv", C.c_float * 3), ("vi", C.c_float * 3), ("ii", C.c_float * 3), ("iv", C.c_float * 3),
                ("pml", C.c_float),
                ("pml_vv", C.c_float * 3), ("pml_vvfn", C.c_float * 3), ("pml_vvfo", C.c_float * 3),
                ("pml_ii", C.c_float * 3), ("pml_iifn", C.c_float * 3), ("pml_iifo", C.c_float * 3),
                ("reserved", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("n_unique", C.c_uint), ("index_bytes", C.c_int), ("hbm_bytes", C.c_uint64),
                ("kernels_launched", C.c_uint64), ("kernels_per_step", C.c_uint),
                ("pml_cells_lo", C.c_uint), ("pml_cells_hi", C.c_uint), ("uses_graph", C.c_int)]


_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)
_ip = C.POINTER(C.c_int)
_u3 = C.c_uint * 3
_i3 = C.c_int * 3
_vp = C.c_void_p

# every symbol include/openems_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "oems_cuda_abi_version": (C.c_int, []),
    "oems_cuda_last_error": (C.c_char_p, [_vp]),
    "oems_cuda_create": (C.c_int, [C.c_uint, C.c_uint, C.c_uint, C.c_int, C.POINTER(_vp)]),
    "oems_cuda_destroy": (C.c_int, [_vp]),
    "oems_cuda_set_slab": (C.c_int, [_vp, C.c_uint, C.c_uint]),
    "oems_cuda_set_operator_dense": (C.c_int, [_vp, _fp, _fp, _fp, _fp]),
    "oems_cuda_set_operator_compressed": (C.c_int, [_vp, C.c_uint, _vp, _vp, C.c_int]),
    "oems_cuda_set_operator_planes": (C.c_int, [_vp, C.c_uint, _vp, C.c_uint, _vp, _up, C.c_int]),
    "oems_cuda_set_signal": (C.c_int, [_vp, _fp, _fp, C.c_uint, C.c_uint]),
    "oems_cuda_add_excitation": (C.c_int, [_vp, C.c_int, C.c_uint, _up, _up, _fp, _up]),
    "oems_cuda_add_upml": (C.c_int, [_vp, _u3, _u3] + [_fp] * 6),
    "oems_cuda_add_mur": (C.c_int, [_vp, C.c_int, C.c_uint, C.c_uint, C.c_uint * 2, _fp, _fp, C.c_uint]),
    "oems_cuda_add_lorentz": (C.c_int, [_vp, C.c_uint, _up] + [_fp] * 6),
    "oems_cuda_add_rlc": (C.c_int, [_vp, C.c_uint, _ip, _up] + [_fp] * 9),
    "oems_cuda_add_steadystate": (C.c_int, [_vp, C.c_uint, C.c_uint, _up, _up]),
    "oems_cuda_steadystate_check": (C.c_int, [_vp, _dp, _up]),
    "oems_cuda_set_tfsf": (C.c_int, [_vp, _up, _up, C.POINTER(C.c_int), C.POINTER(_up), C.POINTER(_fp), C.POINTER(_fp), C.POINTER(_up), C.POINTER(_fp), C.POINTER(_fp)]),
    "oems_cuda_add_absorbing_sheet": (C.c_int, [_vp, C.c_int, _up, _up, C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "oems_cuda_finalize": (C.c_int, [_vp]),
    "oems_cuda_iterate": (C.c_int, [_vp, C.c_uint]),
    "oems_cuda_iterate_timed": (C.c_int, [_vp, C.c_uint, _dp]),
    "oems_cuda_sync": (C.c_int, [_vp]),
    "oems_cuda_num_ts": (C.c_int, [_vp, _up]),
    "oems_cuda_reset": (C.c_int, [_vp]),
    "oems_cuda_add_probe_voltage": (C.c_int, [_vp, _u3, _u3, _ip]),
    "oems_cuda_add_probe_current": (C.c_int, [_vp, _u3, _u3, C.c_int, _i3, _i3, _ip]),
    "oems_cuda_add_probe_field": (C.c_int, [_vp, C.c_int, _u3, _ip]),
    "oems_cuda_num_probe_values": (C.c_int, [_vp, _up]),
    "oems_cuda_read_probes": (C.c_int, [_vp, _dp]),
    "oems_cuda_record_probes": (C.c_int, [_vp, C.c_uint, C.c_uint]),
    "oems_cuda_read_probe_series": (C.c_int, [_vp, _dp, _up, C.c_uint, _up]),
    "oems_cuda_energy": (C.c_int, [_vp, _dp]),
    "oems_cuda_add_dump": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, _up, _up, _up,
                                     C.POINTER(_dp), C.POINTER(_dp), _ip]),
    "oems_cuda_read_dump": (C.c_int, [_vp, C.c_int, _fp]),
    "oems_cuda_fill_fields": (C.c_int, [_vp, C.c_ulonglong]),
    "oems_cuda_field_digest": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_ulonglong)]),
    "oems_cuda_read_dump_async": (C.c_int, [_vp, C.c_int, C.c_void_p, C.POINTER(C.c_longlong)]),
    "oems_cuda_wait": (C.c_int, [_vp, C.c_longlong]),
    "oems_cuda_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "oems_cuda_host_free": (C.c_int, [C.c_void_p]),
    "oems_cuda_get_field": (C.c_int, [_vp, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, _fp]),
    "oems_cuda_set_field": (C.c_int, [_vp, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_float]),
    "oems_cuda_get_fields": (C.c_int, [_vp, C.c_int, _fp]),
    "oems_cuda_set_fields": (C.c_int, [_vp, C.c_int, _fp]),
    "oems_cuda_get_upml_flux": (C.c_int, [_vp, C.c_int, C.c_int, _fp]),
    "oems_cuda_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "oems_cuda_set_tuning": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "oems_cuda_add_fd_dump": (C.c_int, [_vp, C.c_int, C.c_uint, C.POINTER(C.c_int)]),
    "oems_cuda_fd_accumulate": (C.c_int, [_vp, C.c_int, _fp]),
    "oems_cuda_read_fd": (C.c_int, [_vp, C.c_int, _fp, _up]),
    "oems_cuda_add_mode_match": (C.c_int, [_vp, C.c_int, C.c_int, _up, _up, _dp, _dp, _dp, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(C.c_int)]),
    "oems_cuda_read_mode_match": (C.c_int, [_vp, C.c_int, _dp]),
    "oems_cuda_read_mode_match_raw": (C.c_int, [_vp, C.c_int, _dp]),
    "oems_cuda_exchange_ghosts": (C.c_int, [_vp]),
    "oems_cuda_release_ghosts": (C.c_int, [_vp]),
    "oems_cuda_dump_own_range": (C.c_int, [_vp, C.c_int, _up, _up]),
    "oems_cuda_steadystate_raw": (C.c_int, [_vp, _up, _dp, _dp, C.c_uint, _up]),
    "oems_cuda_steadystate_eval": (C.c_int, [C.c_uint, C.c_uint, _up, _dp, _dp, _dp]),
    "oems_cuda_get_option": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_longlong)]),
    "oems_cuda_set_option": (C.c_int, [_vp, C.c_char_p, C.c_longlong]),
    "oems_cuda_time_schedule": (C.c_int, [_vp, C.c_uint, _dp, C.c_uint, _up]),
    "oems_cuda_schedule_label": (C.c_char_p, [_vp, C.c_uint]),
    "oems_cuda_export_ipc": (C.c_int, [_vp, C.POINTER(C.c_ubyte)]),
    "oems_cuda_open_peers": (C.c_int, [_vp, C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte)]),
    "oems_cuda_link_peers": (C.c_int, [_vp, _vp, _vp]),
    # host-side synthetic operator builder (csrc/host/synthetic_operator.h)
    "oems_synth_create": (_vp, [C.c_uint, C.c_uint, C.c_uint, _dp, _dp, _dp, C.c_double]),
    "oems_synth_destroy": (None, [_vp]),
    "oems_synth_set_bc": (None, [_vp, C.c_int * 6, C.c_uint * 6]),
    "oems_synth_set_background": (None, [_vp] + [C.c_double] * 4),
    "oems_synth_set_timestep": (None, [_vp, C.c_double, C.c_double]),
    "oems_synth_add_material": (C.c_int, [_vp, C.c_int, C.c_double * 3, C.c_double * 3] + [C.c_double] * 4),
    "oems_synth_add_metal": (C.c_int, [_vp, C.c_int, C.c_double * 3, C.c_double * 3]),
    "oems_synth_add_excitation": (C.c_int, [_vp, C.c_int, C.c_double * 3, C.c_double * 3, C.c_int, C.c_double * 3, C.c_double]),
    "oems_synth_add_lorentz": (C.c_int, [_vp, C.c_int, C.c_double * 3, C.c_double * 3] + [C.c_double] * 4 + [C.c_int] + [_dp] * 6),
    "oems_synth_set_excite_gauss": (None, [_vp, C.c_double, C.c_double]),
    "oems_synth_set_excite_sinus": (None, [_vp, C.c_double]),
    "oems_synth_build": (C.c_int, [_vp, C.c_uint]),
    "oems_synth_set_slab": (C.c_int, [_vp, C.c_uint, C.c_uint]),
    "oems_synth_local_timestep": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "oems_synth_last_error": (C.c_char_p, [_vp]),
    "oems_synth_dT": (C.c_double, [_vp]),
    "oems_synth_nyquist": (C.c_uint, [_vp]),
    "oems_synth_n_unique": (C.c_uint, [_vp]),
    "oems_synth_index_bytes": (C.c_int, [_vp]),
    "oems_synth_table": (_vp, [_vp]),
    "oems_synth_index": (_vp, [_vp]),
    "oems_synth_unique_planes": (C.c_uint, [_vp]),
    "oems_synth_signal_length": (C.c_uint, [_vp]),
    "oems_synth_signal": (_fp, [_vp, C.c_int]),
    "oems_synth_exc_count": (C.c_uint, [_vp, C.c_int]),
    "oems_synth_exc_get": (None, [_vp, C.c_int, _up, _up, _fp, _up]),
    "oems_synth_mur_count": (C.c_int, [_vp]),
    "oems_synth_mur_coeff": (_fp, [_vp, C.c_int, C.c_int, _ip, _up, _up, C.c_uint * 2, _up]),
    "oems_synth_upml_count": (C.c_int, [_vp]),
    "oems_synth_upml_box": (None, [_vp, C.c_int, _u3, _u3]),
    "oems_synth_lorentz_order": (C.c_int, [_vp]),
    "oems_synth_lorentz_count": (C.c_uint, [_vp, C.c_int]),
    "oems_synth_upload": (C.c_int, [_vp, _vp]),
    "oems_synth_pin": (C.c_int, [_vp]),
    "oems_synth_plane_of_z": (_up, [_vp]),
    "oems_synth_plane_data": (_vp, [_vp]),
}

_lib = None


def library_path():
    return _PATH


def load_library():
    """dlopen libopenems_b200.so; raises LibraryNotBuilt if it is missing or lacks a symbol"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_PATH):
        raise LibraryNotBuilt(
            "libopenems_b200.so is not built (%s). Run `python -m openems_b200.build` or "
            "__graft_entry__.build(); there is no CPU fallback." % _PATH)
    L = C.CDLL(_PATH)
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            f = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        f.restype = res
        f.argtypes = args
    if missing:
        raise LibraryNotBuilt("libopenems_b200.so lacks symbols: " + ", ".join(missing))
    if L.oems_cuda_abi_version() != 1:
        raise LibraryNotBuilt("libopenems_b200.so ABI version mismatch")
    _lib = L
    return L
