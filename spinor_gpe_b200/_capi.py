"""ctypes prototypes of the C ABI declared in include/sgpe.h (one place, shared by the product loader
``_lib.py`` and by the CPU-emulation test harness)."""
import ctypes as C

c_plan = C.c_void_p
c_stream = C.c_void_p
c_dptr = C.c_void_p            # device (or, in the emulation build, host) pointer passed as an integer

PROTOTYPES = {
    'sgpe_last_error': (C.c_char_p, []),
    'sgpe_version': (C.c_char_p, []),
    'sgpe_plan_create': (C.c_int, [C.POINTER(c_plan), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgpe_plan_destroy': (C.c_int, [c_plan]),
    'sgpe_set_grid': (C.c_int, [c_plan, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    'sgpe_set_interactions': (C.c_int, [c_plan, C.c_double, C.c_double, C.c_double]),
    'sgpe_set_kinetic': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int64]),
    'sgpe_set_potential': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int64]),
    'sgpe_set_kinetic_separable': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int64, C.c_int64]),
    'sgpe_set_potential_separable': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int64, C.c_int64]),
    'sgpe_set_coupling': (C.c_int, [c_plan, C.c_int, c_dptr, C.c_int64, c_dptr, c_dptr]),
    'sgpe_set_energy_coupling': (C.c_int, [c_plan, C.c_int, c_dptr, C.c_int64, c_dptr]),
    'sgpe_set_option': (C.c_int, [c_plan, C.c_char_p, C.c_int]),
    'sgpe_set_time': (C.c_int, [c_plan, C.c_int, C.c_double]),
    'sgpe_load_psik': (C.c_int, [c_plan, c_dptr, c_stream]),
    'sgpe_store_psik': (C.c_int, [c_plan, c_dptr, c_stream]),
    'sgpe_full_steps': (C.c_int, [c_plan, C.c_int, c_dptr, C.c_int64, C.c_int, c_stream]),
    'sgpe_full_steps_energy': (C.c_int, [c_plan, C.c_int, c_dptr, C.c_int64, C.c_int, c_dptr, C.c_int64, C.c_int, C.c_int,
                                         C.c_double, c_stream]),
    'sgpe_single_step': (C.c_int, [c_plan, C.c_double, c_stream]),
    'sgpe_substeps': (C.c_int, [c_plan, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    'sgpe_fft2d': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int, c_stream]),
    'sgpe_fft1d': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int, C.c_int, c_stream]),
    'sgpe_sumsq': (C.c_int, [c_plan, c_dptr, c_dptr, c_stream]),
    'sgpe_normalise': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_double, c_stream]),
    'sgpe_energy': (C.c_int, [c_plan, c_dptr, C.c_int, C.c_double, c_dptr, c_stream]),
    'sgpe_kinetic_spectral': (C.c_int, [c_plan, c_dptr, c_dptr, c_stream]),
    'sgpe_gradient': (C.c_int, [c_plan, c_dptr, C.c_int, C.c_double, C.c_double, c_dptr, c_dptr, c_stream]),
    'sgpe_energy_real_space': (C.c_int, [c_plan, c_dptr, C.c_int, C.c_double, c_dptr, c_stream]),
    'sgpe_unwrap_phase': (C.c_int, [c_plan, c_dptr, C.c_int, C.c_int, C.c_int, c_dptr, c_stream]),
    'sgpe_plan_create_lines': (C.c_int, [C.POINTER(c_plan), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgpe_pass_mid': (C.c_int, [c_plan, c_dptr, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_dptr, C.c_double,
                                C.c_int, C.c_int, c_stream]),
    'sgpe_pass_rows': (C.c_int, [c_plan, c_dptr, C.c_double, c_dptr, C.c_double, C.c_int, c_stream]),
    'sgpe_pass_klines': (C.c_int, [c_plan, c_dptr, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, c_dptr,
                                   C.c_int, c_stream]),
    'sgpe_pass_kcols': (C.c_int, [c_plan, c_dptr, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, c_dptr,
                                  C.c_int, c_stream]),
    'sgpe_slab_set_peers': (C.c_int, [c_plan, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                      C.c_int]),
    'sgpe_slab_window': (C.c_int, [c_plan, C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgpe_ipc_alloc': (C.c_int, [C.c_int, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]),
    'sgpe_ipc_open': (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    'sgpe_ipc_close': (C.c_int, [C.c_void_p]),
    'sgpe_ipc_free': (C.c_int, [C.c_void_p]),
    'sgpe_slab_pack': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int, C.c_int, C.c_int, c_stream]),
    'sgpe_slab_unpack': (C.c_int, [c_plan, c_dptr, c_dptr, C.c_int, C.c_int, C.c_int, c_stream]),
    'sgpe_run_host': (C.c_int, [c_plan, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, c_stream]),
    'sgpe_step_accounting': (C.c_int, [c_plan, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    'sgpe_profile_begin': (C.c_int, [c_plan]),
    'sgpe_profile_end': (C.c_int, [c_plan, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                   C.POINTER(C.c_uint64)]),
    'sgpe_debug_timeline': (C.c_int, [c_plan, c_dptr]),
    'sgpe_launch_count': (C.c_int, [c_plan, C.POINTER(C.c_uint64)]),
}

SGPE_C128, SGPE_C64 = 0, 1
SGPE_TIME_REAL, SGPE_TIME_IMAG = 0, 1
SGPE_COUPLING_NONE, SGPE_COUPLING_UNIFORM, SGPE_COUPLING_DENSE = 0, 1, 2


class SgpeError(RuntimeError):
    """A C-ABI call returned a negative status."""


def bind(lib):
    """Attach argtypes / restype for every exported symbol; raises AttributeError if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def check(lib, rc, what=''):
    if rc != 0:
        msg = lib.sgpe_last_error()
        raise SgpeError(f"{what or 'sgpe call'} failed ({rc}): {msg.decode() if msg else '?'}")
