"""ctypes loader for libmdns_b200.so (the C ABI declared in include/mdns_b200.h).

The library is built in-tree by ``make -C massivedatans_b200/csrc`` (or
``__graft_entry__.build()``).  There is no fallback: a missing library, a missing
symbol or a failing call raises.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint8, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmdns_b200.so')
DROPIN_DIR = os.path.join(_HERE, 'dropin')

MDNS_OK = 0

_P = c_void_p   # opaque handles and raw array addresses

# name -> (restype, argtypes); mirrors include/mdns_b200.h one to one
SIGNATURES = {
    'mdns_last_error': (c_char_p, []),
    'mdns_version': (c_int, []),
    'mdns_device_count': (c_int, []),
    'mdns_launch_count': (c_int64, []),
    'mdns_last_kernel': (c_char_p, []),
    'mdns_sqrt_threshold': (c_double, [c_double]),
    'mdns_host_alloc': (c_void_p, [c_int64]),
    'mdns_host_free': (c_int, [c_void_p]),
    'mdns_dataset_create': (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, POINTER(c_void_p)]),
    'mdns_dataset_create_from_npy': (c_int, [_P, c_char_p, c_char_p, _P, c_int, POINTER(c_void_p)]),
    'mdns_dataset_destroy': (c_int, [_P]),
    'mdns_dataset_info': (c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                  POINTER(c_int64)]),
    'mdns_clike_eval_params': (c_int, [_P, _P, c_int, c_double, c_double, _P, _P, c_int64,
                                       POINTER(c_int)]),
    'mdns_clike_eval_spectra': (c_int, [_P, _P, c_int, c_double, c_double, _P, _P, c_int64,
                                        POINTER(c_int)]),
    'mdns_muse_eval_spectra': (c_int, [_P, _P, c_int, _P, _P]),
    'mdns_set_mask': (c_int, [_P, _P, POINTER(c_int)]),
    'mdns_stage_params': (c_int, [_P, _P, c_int]),
    'mdns_stage_spectra': (c_int, [_P, _P, c_int]),
    'mdns_clike_launch': (c_int, [_P, c_double, c_double]),
    'mdns_muse_launch': (c_int, [_P]),
    'mdns_clike_launch_fetch': (c_int, [_P, c_double, c_double, _P, c_int64]),
    'mdns_set_thresholds': (c_int, [_P, _P]),
    'mdns_clike_first_accept': (c_int, [_P, c_double, c_double, _P, _P, POINTER(c_int), _P,
                                        c_int64]),
    'mdns_clike_draw_pass': (c_int, [_P, _P, _P, _P, c_int, c_double, c_double, _P, POINTER(c_int), _P,
                                     c_int64, POINTER(c_int)]),
    'mdns_clike_first_accept_sparse': (c_int, [_P, c_double, c_double, _P, _P, POINTER(c_int), _P, _P,
                                               c_int64, POINTER(c_int)]),
    'mdns_clike_accept_counts': (c_int, [_P, c_double, c_double, _P, _P]),
    'mdns_fetch_candidate': (c_int, [_P, c_int, _P, c_int64]),
    'mdns_set_expanded': (c_int, [_P, c_int, c_double]),
    'mdns_expanded_stats': (c_int, [_P, POINTER(c_int), POINTER(c_int64)]),
    'mdns_fetch': (c_int, [_P, _P, c_int64]),
    'mdns_livetable_create': (c_int, [_P, c_int, POINTER(_P)]),
    'mdns_livetable_destroy': (c_int, [_P]),
    'mdns_livetable_upload': (c_int, [_P, _P]),
    'mdns_livetable_download': (c_int, [_P, _P]),
    'mdns_livetable_fill_from_launch': (c_int, [_P, _P, c_int]),
    'mdns_livetable_colstats': (c_int, [_P, _P, _P, _P]),
    'mdns_livetable_replace': (c_int, [_P, _P, _P]),
    'mdns_livetable_stage_thresholds': (c_int, [_P, _P]),
    'mdns_livetable_lmins_higher': (c_int, [_P, _P, c_int, _P, _P, _P]),
    'mdns_livetable_upload_points': (c_int, [_P, _P]),
    'mdns_livetable_replace_points': (c_int, [_P, _P, _P]),
    'mdns_livetable_subsets': (c_int, [_P, _P, c_int64, _P, POINTER(c_int), POINTER(c_int)]),
    'mdns_sync': (c_int, [_P]),
    'mdns_set_draw_chunks': (c_int, [_P, c_int]),
    'mdns_comm_unique_id': (c_int, [_P]),
    'mdns_comm_init': (c_int, [_P, _P, c_int, c_int]),
    'mdns_comm_destroy': (c_int, [_P]),
    'mdns_comm_info': (c_int, [_P, POINTER(c_int), POINTER(c_int)]),
    'mdns_comm_allreduce': (c_int, [_P, _P, c_int, c_int]),
    'mdns_comm_allgather_candidate': (c_int, [_P, c_int, _P, c_int64, _P]),
    'mdns_timer_start': (c_int, [_P]),
    'mdns_timer_stop': (c_int, [_P, POINTER(c_float)]),
    'mdns_flush_l2': (c_int, [_P]),
    'mdns_region_timer_start': (c_int, [_P]),
    'mdns_region_timer_stop': (c_int, [_P, POINTER(c_float)]),
    'mdns_set_tuning': (c_int, [_P, c_int, c_int, c_int, c_int]),
    'mdns_muse_model_create': (c_int, [_P, _P, c_int, _P, c_int, _P, c_int, _P, _P, _P, c_int, c_int,
                                       POINTER(_P)]),
    'mdns_muse_model_destroy': (c_int, [_P]),
    'mdns_muse_model_stage': (c_int, [_P, _P, c_int, _P]),
    'mdns_muse_model_spectra': (c_int, [_P, _P]),
    'mdns_region_create': (c_int, [c_int, POINTER(c_void_p)]),
    'mdns_region_destroy': (c_int, [_P]),
    'mdns_region_set_members': (c_int, [_P, _P, c_int, c_int]),
    'mdns_region_count_within': (c_int, [_P, c_double, _P, c_int, _P, c_int]),
    'mdns_region_generate': (c_int, [_P, c_double, ctypes.c_uint64, ctypes.c_uint64, c_int, _P, c_int64,
                                     POINTER(c_int)]),
    'mdns_region_nearest_index': (c_int, [_P, _P]),
    'mdns_region_axis_covered': (c_int, [_P, _P, _P, c_int, _P, c_int, _P]),
    'mdns_region_is_within': (c_int, [_P, c_double, _P, POINTER(c_int)]),
    'mdns_region_bootstrapped_maxdistance': (c_int, [_P, _P, c_int, POINTER(c_double)]),
    'mdns_region_most_distant_nearest_neighbor': (c_int, [_P, POINTER(c_double)]),
    'mdns_most_distant_nearest_neighbor': (c_double, [_P, c_int, c_int]),
    'mdns_is_within_distance_of': (c_int, [_P, c_int, c_int, c_double, _P]),
    'mdns_count_within_distance_of': (c_int, [_P, c_int, c_int, c_double, _P, c_int, _P, c_int]),
    'mdns_bootstrapped_maxdistance': (c_double, [_P, c_int, c_int, _P, c_int]),
    'mdns_clike_like': (c_int, [_P, _P, c_int, c_int, c_double, c_double, c_double, c_double,
                                _P, _P]),
    'mdns_cmuselike_like': (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    'mdns_legacy_reset': (c_int, []),
    'mdns_legacy_trust': (c_int, [c_int]),
}

_lib = None


class MdnsError(RuntimeError):
    pass


def load():
    """Load libmdns_b200.so and bind every declared entry point (raises if absent)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MdnsError('%s not built: run `make -C massivedatans_b200/csrc` '
                            '(there is no CPU fallback)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error():
    return load().mdns_last_error().decode('utf-8', 'replace')


def check(rc, what):
    if rc != MDNS_OK:
        raise MdnsError('%s failed (%d): %s' % (what, rc, last_error()))


def require_device():
    n = load().mdns_device_count()
    if n <= 0:
        raise MdnsError('no CUDA device visible: massivedatans_b200 has no CPU fallback')
    return n


__all__ = ['load', 'check', 'last_error', 'require_device', 'MdnsError', 'SIGNATURES',
           'LIB_PATH', 'DROPIN_DIR', 'c_uint8']
