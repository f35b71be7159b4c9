"""ctypes loader of libftb200.so (the C-ABI in include/ftb200.h).

There is no CPU fallback: if the library is missing it is NOT silently
replaced by anything -- importing callers get a hard error telling them to
build it (python -m femtech_b200.build), and creating a context without a CUDA
device fails in ftb200_create().
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FTB200_LIB selects an alternative build of the same library (kernel tuning experiments only)
LIB_PATH = os.environ.get("FTB200_LIB") or os.path.join(HERE, "libftb200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_ll = C.c_longlong

# name -> (restype, argtypes); must cover every symbol declared in include/ftb200.h
SIGNATURES = {
    "ftb200_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ftb200_destroy": (C.c_int, [_vp]),
    "ftb200_last_error": (C.c_char_p, [_vp]),
    "ftb200_set_stream": (C.c_int, [_vp, _vp]),
    "ftb200_launch_count": (_ll, [_vp]),
    "ftb200_build_info": (C.c_char_p, []),
    "ftb200_upload_mesh": (C.c_int, [_vp, _dp, _ip, _ip, C.c_int, C.c_int]),
    "ftb200_upload_materials": (C.c_int, [_vp, _ip, _dp, C.c_int]),
    "ftb200_upload_comm": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip]),
    "ftb200_shape_functions": (C.c_int, [_vp, _dp]),
    "ftb200_lumped_mass": (C.c_int, [_vp, _dp]),
    "ftb200_get_mass": (C.c_int, [_vp, _dp]),
    "ftb200_get_force": (C.c_int, [_vp, _dp, _dp, C.c_double, _dp, _dp]),
    "ftb200_calculate_accelerations": (C.c_int, [_vp, _ip, _dp]),
    "ftb200_stable_time_step": (C.c_int, [_vp, _dp, _ip, _dp]),
    "ftb200_check_energy": (C.c_int, [_vp] + [_dp] * 9 + [_ip, _dp]),
    "ftb200_get_gp_outputs": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "ftb200_set_state": (C.c_int, [_vp, _dp, _dp, _dp, _ip]),
    "ftb200_get_state": (C.c_int, [_vp, _dp, _dp, _dp, _ip, _dp, _dp]),
    "ftb200_set_bc": (C.c_int, [_vp, _ip, _dp]),
    "ftb200_explicit_begin": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_int]),
    "ftb200_explicit_run": (C.c_int, [_vp, C.c_double, _ll, C.POINTER(_ll), _dp, _dp]),
    "ftb200_explicit_run_async": (C.c_int, [_vp, C.c_double, _ll]),
    "ftb200_explicit_poll": (C.c_int, [_vp, C.POINTER(_ll), _dp, _dp, _ip]),
    "ftb200_step_ring": (C.c_int, [_vp, _ll, C.POINTER(_dp)]),
    "ftb200_get_energy": (C.c_int, [_vp, _dp]),
    "ftb200_record_history": (C.c_int, [_vp, _ll]),
    "ftb200_get_history": (C.c_int, [_vp, _ll, _ll, _dp, _dp]),
    "ftb200_halo_count": (C.c_int, [_vp]),
    "ftb200_halo_pack": (C.c_int, [_vp, C.c_int, _vp]),
    "ftb200_halo_add": (C.c_int, [_vp, C.c_int, _vp]),
    "ftb200_run_begin": (C.c_int, [_vp, C.c_double, _ll]),
    "ftb200_step_begin": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "ftb200_step_join": (C.c_int, [_vp]),
    "ftb200_step_end": (C.c_int, [_vp, _vp]),
    "ftb200_explicit_begin_dt": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(_vp)]),
    "ftb200_explicit_begin_force": (C.c_int, [_vp, _vp]),
    "ftb200_explicit_begin_finish": (C.c_int, [_vp, _vp]),
    "ftb200_p2p_export": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "ftb200_p2p_import": (C.c_int, [_vp, _vp, C.c_int, _ip, _ip, _ip]),
    "ftb200_profile_enable": (C.c_int, [_vp, C.c_int]),
    "ftb200_profile_get": (C.c_int, [_vp, _dp, _dp, C.POINTER(_ll), C.POINTER(_ll)]),
    "ftb200_measure_peaks": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "ftb200_upload_mesh_mixed": (C.c_int, [_vp, _dp, _ip, _ip, _ip, C.c_int, C.c_int]),
    "ftb200_gauss_point_count": (_ll, [_vp]),
    "ftb200_affine_element_count": (_ll, [_vp]),
    "ftb200_brick_info": (C.c_int, [_vp, C.POINTER(_ll)]),
    "ftb200_device_count": (C.c_int, []),
    "ftb200_halo_pack_host": (C.c_int, [_vp, C.c_int, _dp]),
    "ftb200_halo_add_host": (C.c_int, [_vp, C.c_int, _dp]),
    "ftb200_get_force_begin": (C.c_int, [_vp, _dp, _dp, C.c_double, _dp]),
    "ftb200_get_force_end": (C.c_int, [_vp, _dp, _dp, _dp]),
    "ftb200_get_dtmin": (C.c_int, [_vp, _dp]),
    "ftb200_set_dtmin": (C.c_int, [_vp, C.c_double]),
    "ftb200_explicit_begin_force_host": (C.c_int, [_vp, _dp]),
    "ftb200_explicit_begin_finish_host": (C.c_int, [_vp, _dp]),
    "ftb200_step_begin_host": (C.c_int, [_vp, _dp, _dp]),
    "ftb200_step_end_host": (C.c_int, [_vp, _dp, C.c_double]),
    "ftb200_brick_maps": (C.c_int, [_vp, _ip, _ip]),
    "ftb200_set_rigid_bc": (C.c_int, [_vp, _ip, C.POINTER(_dp), C.POINTER(_dp), _ip, C.c_int]),
    "ftb200_get_rigid_state": (C.c_int, [_vp, _dp, _dp, _ip]),
    "ftb200_explicit_poll_async": (C.c_int, [_vp, _dp]),
    "ftb200_injury_begin": (C.c_int, [_vp, _ip, C.c_int, _dp]),
    "ftb200_injury_end": (C.c_int, [_vp]),
    "ftb200_injury_local_count": (C.c_int, [_vp, C.POINTER(_ll)]),
    "ftb200_injury_global_count": (C.c_int, [_vp, _ll]),
    "ftb200_injury_passes": (C.c_int, []),
    "ftb200_injury_select_hist": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), _ip]),
    "ftb200_injury_select_pick": (C.c_int, [_vp, C.c_int]),
    "ftb200_injury_get": (C.c_int, [_vp, _dp, _ip, C.POINTER(C.c_ubyte), _dp, _dp, _dp]),
    "ftb200_injury_history": (C.c_int, [_vp, _ll, _ll, _dp, _dp]),
    "ftb200_principal_strains": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
}

_lib = None


class FemTechB200Error(RuntimeError):
    """A C-ABI call failed; .code is the reference's TerminateFemTech code."""

    def __init__(self, code, msg):
        super().__init__("ftb200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load libftb200.so and bind every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build the CUDA library with `python -m femtech_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
