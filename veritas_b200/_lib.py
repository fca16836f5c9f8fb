"""ctypes binding of include/veritas_b200.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libveritas_b200.so")
_lib = None
dbl_p = C.POINTER(C.c_double)


class VrtError(RuntimeError):
    pass


class PatchDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("depth", "x_pos", "p_pos", "n_x", "n_p", "up", "down", "left", "right")]


class CaseParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("lambda_", "a0", "density", "temp_frac", "pmax_e", "pmax_i", "box_lambdas", "ion_mass_ratio")]


class CaseDerived(C.Structure):
    _fields_ = [("dp", C.c_double * 2), ("pmin", C.c_double * 2), ("dx", C.c_double), ("sizeWeight", C.c_double),
                ("temp0", C.c_double * 2), ("temp1", C.c_double * 2), ("tempEM", C.c_double * 2), ("quadratureDepth", C.c_int)]


# name -> (restype, argtypes); every symbol declared in include/veritas_b200.h
SIGNATURES = {
    "vrt_global_error": (C.c_char_p, []),
    "vrt_last_error": (C.c_char_p, [C.c_void_p]),
    "vrt_version": (C.c_char_p, []),
    "vrt_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "vrt_destroy": (C.c_int, [C.c_void_p]),
    "vrt_sync": (C.c_int, [C.c_void_p]),
    "vrt_stream": (C.c_void_p, [C.c_void_p]),
    "vrt_set_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vrt_set_species": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "vrt_set_hierarchy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(PatchDesc)]),
    "vrt_regrid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(PatchDesc)]),
    "vrt_patch_energy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, dbl_p]),
    "vrt_get_hierarchy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(PatchDesc)]),
    "vrt_checkpoint_write": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vrt_checkpoint_read": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vrt_error_flags": (C.c_int, [C.c_void_p, C.c_int, C.c_int, dbl_p, C.c_double, C.POINTER(C.c_ubyte)]),
    "vrt_set_path": (C.c_int, [C.c_void_p, C.c_int]),
    "vrt_get_path": (C.c_int, [C.c_void_p, C.c_int]),
    "vrt_patch_upload_f": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, dbl_p]),
    "vrt_patch_download_f": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, dbl_p]),
    "vrt_patch_download_plane": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dbl_p]),
    "vrt_commit_state": (C.c_int, [C.c_void_p, C.c_int]),
    "vrt_conn_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(PatchDesc), C.c_int, C.c_int]),
    "vrt_conn_strips": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_ubyte)]),
    "vrt_conn_flags": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_ubyte)]),
    "vrt_conn_destroy": (None, [C.c_void_p]),
    "vrt_field_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int, dbl_p]),
    "vrt_field_download": (C.c_int, [C.c_void_p, C.c_int, C.c_int, dbl_p]),
    "vrt_set_1d": (C.c_int, [C.c_void_p, C.c_int, dbl_p]),
    "vrt_get_1d": (C.c_int, [C.c_void_p, C.c_int, dbl_p]),
    "vrt_set_scalar": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "vrt_get_scalar": (C.c_int, [C.c_void_p, C.c_int, dbl_p]),
    "vrt_moments": (C.c_int, [C.c_void_p]),
    "vrt_enforce_neutralization": (C.c_int, [C.c_void_p]),
    "vrt_poisson": (C.c_int, [C.c_void_p]),
    "vrt_vlasov_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "vrt_vlasov_substep": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]),
    "vrt_push_data": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vrt_push_boundary_c": (C.c_int, [C.c_void_p, C.c_int]),
    "vrt_level_push": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vrt_moments_species": (C.c_int, [C.c_void_p, C.c_int, dbl_p, dbl_p]),
    "vrt_patch_moments": (C.c_int, [C.c_void_p, C.c_int, C.c_int, dbl_p, dbl_p]),
    "vrt_field_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]),
    "vrt_cfl_bound": (C.c_int, [C.c_void_p, dbl_p]),
    "vrt_update_time": (C.c_double, [C.c_double, C.c_int, C.c_double]),
    "vrt_step": (C.c_int, [C.c_void_p, C.c_double, dbl_p]),
    "vrt_step_fields": (C.c_int, [C.c_void_p, C.c_double, dbl_p]),
    "vrt_last_step_launches": (C.c_long, [C.c_void_p]),
    "vrt_fused_plan": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "vrt_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vrt_init_maxwellian_slab": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "vrt_set_slab": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vrt_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "vrt_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "vrt_case_derive": (C.c_int, [C.POINTER(CaseParams), C.c_double, C.c_double, C.c_uint, C.POINTER(C.c_uint), C.c_double, C.POINTER(CaseDerived)]),
    "vrt_case_laser_by": (C.c_double, [C.c_double] * 4),
    "vrt_case_laser_bz": (C.c_double, [C.c_double] * 4),
    "vrt_case_maxwellian_slab": (C.c_double, [C.c_double] * 6),
}


def load():
    """Load the CUDA library.  Fails loudly if it has not been built (python -m veritas_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VrtError(f"{LIB_PATH} is missing: build it with `python -m veritas_b200.build` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
