"""ctypes binding of libcu2b.so (include/cu2b.h). Fails loudly when the library is missing:
there is no Python or CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcu2b.so")


class Rating(C.Structure):  # cu2b_rating
    _fields_ = [("user", C.c_int32), ("item", C.c_int32), ("rating", C.c_float)]


class Config(C.Structure):  # cu2b_config
    _fields_ = [
        ("cur_iterations", C.c_int), ("total_iterations", C.c_int), ("n_factors", C.c_int),
        ("learning_rate", C.c_float), ("seed", C.c_int),
        ("P_reg", C.c_float), ("Q_reg", C.c_float), ("user_bias_reg", C.c_float), ("item_bias_reg", C.c_float),
        ("is_train", C.c_int), ("n_threads", C.c_int), ("check_error", C.c_int),
        ("patience", C.c_float), ("learning_rate_decay", C.c_float),
        ("mode", C.c_int), ("sampler", C.c_int), ("n_blocks", C.c_int), ("n_gpus", C.c_int), ("round_iters", C.c_int),
    ]


class Csr(C.Structure):  # cu2b_csr
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("nonzeros", C.c_int),
                ("indptr", C.c_void_p), ("indices", C.c_void_p), ("data", C.c_void_p),
                ("on_device", C.c_int)]


class Metrics(C.Structure):  # cu2b_metrics
    _fields_ = [("iteration", C.c_int), ("train_mae", C.c_float), ("train_rmse", C.c_float),
                ("test_mae", C.c_float), ("test_rmse", C.c_float), ("learning_rate", C.c_float)]


class Stats(C.Structure):  # cu2b_stats
    _fields_ = [("sgd_ms", C.c_double), ("loss_ms", C.c_double), ("sampler_ms", C.c_double),
                ("total_ms", C.c_double), ("updates", C.c_int64), ("kernel_launches", C.c_int64),
                ("sgd_launches", C.c_int64), ("wait_ms", C.c_double), ("send_ms", C.c_double)]


# every symbol include/cu2b.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "cu2b_last_error": (C.c_char_p, []),
    "cu2b_version": (C.c_int, []),
    "cu2b_free": (None, [_P]),
    "cu2b_config_default": (None, [C.POINTER(Config)]),
    "cu2b_config_read": (C.c_int, [C.c_char_p, C.POINTER(Config)]),
    "cu2b_config_write": (C.c_int, [C.c_char_p, C.POINTER(Config)]),
    "cu2b_config_format": (C.c_int, [C.POINTER(Config), C.c_char_p, C.c_int]),
    "cu2b_read_csv": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(Rating)), C.POINTER(C.c_int64),
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "cu2b_read_csv_cached": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.POINTER(Rating)), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "cu2b_build_csr": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, _P]),
    "cu2b_read_array": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cu2b_write_csv": (C.c_int, [C.c_char_p, _P, C.c_int, C.c_int]),
    "cu2b_write_component": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _P, C.c_int, C.c_int, C.c_int]),
    "cu2b_init_normal": (None, [_P, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int]),
    "cu2b_synth_ratings": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_float, C.c_uint64,
                                     _P, C.POINTER(C.c_int64), _P, C.POINTER(C.c_int64)]),
    "cu2b_loss": (C.c_int, [C.POINTER(Csr), _P, _P, _P, _P, C.c_float, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cu2b_residuals": (C.c_int, [C.POINTER(Csr), _P, _P, _P, _P, C.c_float, C.c_int, _P]),
    "cu2b_error_metrics": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cu2b_sample_per_user": (C.c_int, [C.POINTER(Csr), C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int64)]),
    "cu2b_sample_per_rating": (C.c_int, [C.POINTER(Csr), C.c_int, C.c_int64, C.c_int64, _P]),
    "cu2b_sgd_apply": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, C.c_int, _P, _P, C.c_float, C.POINTER(Config), C.c_int]),
    "cu2b_sgd_blocked": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, C.c_int, _P, _P, C.c_float, C.POINTER(Config), C.c_int, C.c_int]),
    "cu2b_block_schedule_order": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
    "cu2b_train": (C.c_int, [C.POINTER(Csr), C.POINTER(Csr), C.POINTER(Config), _P, _P, _P, _P, C.c_float, C.c_int,
                             _P, C.POINTER(Metrics), C.c_int, C.POINTER(C.c_int), C.POINTER(Stats)]),
    "cu2b_session_create": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(Csr), C.POINTER(Csr), C.POINTER(Config),
                                      _P, _P, _P, _P, C.c_float]),
    "cu2b_session_run": (C.c_int, [_P, C.c_int]),
    "cu2b_session_run_download": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "cu2b_session_reload": (C.c_int, [_P, C.POINTER(Csr), C.POINTER(Csr), _P, _P, _P, _P, C.c_float]),
    "cu2b_dsgd_reload": (C.c_int, [_P, C.POINTER(Csr), C.POINTER(Csr), _P, _P, _P, _P, C.c_float]),
    "cu2b_session_eval": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cu2b_session_log": (C.c_int, [_P, C.POINTER(Metrics), C.c_int, C.POINTER(C.c_int)]),
    "cu2b_session_download": (C.c_int, [_P, _P, _P, _P, _P]),
    "cu2b_session_get_config": (C.c_int, [_P, C.POINTER(Config)]),
    "cu2b_session_stats": (C.c_int, [_P, C.POINTER(Stats), C.c_int]),
    "cu2b_session_destroy": (None, [_P]),
    "cu2b_dsgd_partition": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "cu2b_dsgd_extract_strip": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int, _P, C.POINTER(C.c_int64)]),
    "cu2b_dsgd_item_keep": (C.c_int, [C.POINTER(Csr), _P, C.c_int, C.c_float, C.c_int, C.c_double, _P]),
    "cu2b_dsgd_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.POINTER(Csr), C.POINTER(Csr),
                                   C.POINTER(Config), _P, _P, _P, _P, C.c_float, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P]),
    "cu2b_dsgd_connect": (C.c_int, [_P, _P]),
    "cu2b_dsgd_run": (C.c_int, [_P, C.c_int]),
    "cu2b_dsgd_session": (_P, [_P]),
    "cu2b_dsgd_local_sums": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "cu2b_dsgd_destroy": (None, [_P]),
    "cu2b_prep_map": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char, C.c_int, C.c_int, C.c_char_p, C.c_char_p] +
                      [C.POINTER(C.c_int64)] * 6),
    "cu2b_write_ratings_csv": (C.c_int, [C.c_char_p, C.POINTER(Rating), C.c_int64]),
    "cu2b_prep_sort": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "cu2b_prep_split": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_int64, C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64)]),
    "cu2b_prep_convert_to_np": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cu2b_prep_create_config": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int] + [C.c_double] * 4),
    "cu2b_predict_topk": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, C.c_float, C.c_int, C.POINTER(Csr), C.c_int, _P, _P,
                                    C.POINTER(C.c_double)]),
    "cu2b_release_cache": (C.c_int, []),
    "cu2b_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}

DSGD_HANDLE_BYTES = 512
_lib = None


class Cu2bError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("cu2b status %d: %s" % (status, message))
        self.status = status


def load():
    """Loads libcu2b.so (built in-tree by cu2rec_b200/build.py). No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libcu2b.so is missing at %s -- run `python -m cu2rec_b200.build` (or "
            "__graft_entry__.build()). cu2rec_b200 has no CPU / Python fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise Cu2bError(status, load().cu2b_last_error().decode("utf-8", "replace"))
