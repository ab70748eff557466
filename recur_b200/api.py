"""ctypes binding of librecur_b200.so (include/recur-nn.h, include/recur_b200.h).

Loading fails loudly if the library has not been built: there is no Python or
CPU substitute for it.
"""
import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librecur_b200.so")

c_float_p = abi.c_float_p
u8_p = abi.u8_p


class RnnBatchCharStats(C.Structure):
    _fields_ = [("error", C.c_double), ("entropy", C.c_double),
                ("correct", C.c_int64), ("count", C.c_int64)]


class RnnBatchBpttLog(C.Structure):
    _fields_ = [("depth", C.c_int32), ("n_steps", C.c_int32),
                ("scaled_error", C.c_float), ("ih_scale", C.c_float),
                ("min_error_threshold", C.c_float), ("min_error_factor", C.c_float),
                ("cum_error", C.c_float), ("error_sum", C.c_float),
                ("top_error_scaled", C.c_float), ("top_error_raw", C.c_float)]


B200_API_SYMBOLS = [
    "rnn_b200_device_count", "rnn_b200_set_device", "rnn_b200_synchronize",
    "rnn_b200_stream", "rnn_b200_version", "rnn_b200_kernel_launches",
    "rnn_b200_set_engine", "rnn_b200_last_walk_kernel", "rnn_b200_pull", "rnn_b200_push",
    "rnn_b200_profile_enable", "rnn_b200_profile_read", "rnn_b200_profile_class_name",
    "rnn_batch_new", "rnn_batch_delete", "rnn_batch_size", "rnn_batch_advance",
    "rnn_batch_set_inputs", "rnn_batch_set_one_hot", "rnn_batch_opinion",
    "rnn_batch_get_outputs", "rnn_batch_get_hiddens", "rnn_batch_softmax_error",
    "rnn_batch_set_errors", "rnn_batch_calc_deltas", "rnn_batch_calc_deltas_masked",
    "rnn_batch_apply_learning",
    "rnn_batch_char_step", "rnn_batch_text_upload", "rnn_batch_text_train",
    "rnn_batch_text_forward", "rnn_batch_rnnca_frame", "rnn_batch_pull", "rnn_batch_bptt_depths",
    "rnn_batch_bptt_log",
    "rnn_batch_p2p_probe", "rnn_b200_adaptive_downscale",
    "rnn_b200_adaptive_downscale_device", "rnn_mfcc_new", "rnn_mfcc_delete", "rnn_mfcc_extract",
    "rnn_mfcc_extract_device", "rnn_mfcc_tables", "rnn_cells_new", "rnn_cells_new_sharded", "rnn_cells_delete", "rnn_cells_forget", "rnn_cells_rnnca_frame",
    "rnn_cells_rnnca_run", "rnn_cells_get_hidden",
    "rnn_b200_comm_unique_id", "rnn_b200_comm_join", "rnn_b200_comm_leave",
    "rnn_b200_comm_size", "rnn_batch_p2p_export", "rnn_batch_p2p_attach",
]


def _declare_b200(lib):
    P = abi.RecurNN_p
    vp = C.c_void_p
    lib.rnn_b200_device_count.restype = C.c_int
    lib.rnn_b200_device_count.argtypes = []
    lib.rnn_b200_set_device.restype = C.c_int
    lib.rnn_b200_set_device.argtypes = [C.c_int]
    lib.rnn_b200_synchronize.restype = None
    lib.rnn_b200_synchronize.argtypes = []
    lib.rnn_b200_stream.restype = vp
    lib.rnn_b200_stream.argtypes = []
    lib.rnn_b200_version.restype = C.c_char_p
    lib.rnn_b200_version.argtypes = []
    lib.rnn_b200_kernel_launches.restype = C.c_uint64
    lib.rnn_b200_kernel_launches.argtypes = []
    lib.rnn_b200_set_engine.restype = C.c_int
    lib.rnn_b200_set_engine.argtypes = [C.c_int]
    lib.rnn_b200_last_walk_kernel.restype = C.c_char_p
    lib.rnn_b200_last_walk_kernel.argtypes = []
    lib.rnn_b200_profile_enable.restype = None
    lib.rnn_b200_profile_enable.argtypes = [C.c_int]
    lib.rnn_b200_profile_read.restype = C.c_int
    lib.rnn_b200_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    lib.rnn_b200_profile_class_name.restype = C.c_char_p
    lib.rnn_b200_profile_class_name.argtypes = [C.c_int]
    lib.rnn_b200_pull.restype = None
    lib.rnn_b200_pull.argtypes = [P]
    lib.rnn_b200_push.restype = None
    lib.rnn_b200_push.argtypes = [P]
    lib.rnn_batch_new.restype = vp
    lib.rnn_batch_new.argtypes = [abi.RecurNN_pp, C.c_int]
    lib.rnn_batch_delete.restype = None
    lib.rnn_batch_delete.argtypes = [vp]
    lib.rnn_batch_size.restype = C.c_int
    lib.rnn_batch_size.argtypes = [vp]
    lib.rnn_batch_advance.restype = None
    lib.rnn_batch_advance.argtypes = [vp]
    lib.rnn_batch_set_inputs.restype = None
    lib.rnn_batch_set_inputs.argtypes = [vp, c_float_p]
    lib.rnn_batch_set_one_hot.restype = None
    lib.rnn_batch_set_one_hot.argtypes = [vp, u8_p]
    lib.rnn_batch_opinion.restype = None
    lib.rnn_batch_opinion.argtypes = [vp, C.c_float]
    lib.rnn_batch_get_outputs.restype = None
    lib.rnn_batch_get_outputs.argtypes = [vp, c_float_p]
    lib.rnn_batch_get_hiddens.restype = None
    lib.rnn_batch_get_hiddens.argtypes = [vp, c_float_p]
    lib.rnn_batch_softmax_error.restype = None
    lib.rnn_batch_softmax_error.argtypes = [vp, u8_p, c_float_p,
                                            C.POINTER(C.c_int32)]
    lib.rnn_batch_set_errors.restype = None
    lib.rnn_batch_set_errors.argtypes = [vp, c_float_p]
    lib.rnn_batch_calc_deltas.restype = None
    lib.rnn_batch_calc_deltas.argtypes = [vp, C.c_int]
    lib.rnn_batch_calc_deltas_masked.restype = None
    lib.rnn_batch_calc_deltas_masked.argtypes = [vp, C.c_int, u8_p]
    lib.rnn_batch_apply_learning.restype = None
    lib.rnn_batch_apply_learning.argtypes = [vp, C.c_int, C.c_float]
    lib.rnn_batch_char_step.restype = None
    lib.rnn_batch_char_step.argtypes = [vp, u8_p, u8_p, C.c_int, C.c_float,
                                        C.POINTER(RnnBatchCharStats)]
    lib.rnn_batch_text_upload.restype = None
    lib.rnn_batch_text_upload.argtypes = [vp, u8_p, C.c_int]
    lib.rnn_batch_text_train.restype = C.c_int
    lib.rnn_batch_text_train.argtypes = [vp, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_float,
                                         C.POINTER(RnnBatchCharStats)]
    lib.rnn_batch_text_forward.restype = C.c_int
    lib.rnn_batch_text_forward.argtypes = [vp, C.c_int, C.c_int]
    lib.rnn_batch_rnnca_frame.restype = None
    lib.rnn_batch_rnnca_frame.argtypes = [vp, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int,
                                          C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int),
                                          C.c_int, C.c_int, C.c_int]
    lib.rnn_batch_bptt_depths.restype = None
    lib.rnn_batch_bptt_depths.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.rnn_batch_p2p_probe.restype = C.c_float
    lib.rnn_batch_p2p_probe.argtypes = [vp, C.c_int]
    ip, bp = C.POINTER(C.c_int), C.POINTER(C.c_uint8)
    lib.rnn_cells_new.restype = vp
    lib.rnn_cells_new.argtypes = [P, C.c_int, C.c_int]
    for name in ("rnn_b200_adaptive_downscale", "rnn_b200_adaptive_downscale_device"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    lib.rnn_mfcc_new.restype = vp
    lib.rnn_mfcc_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_float, C.c_int]
    lib.rnn_mfcc_delete.restype = None
    lib.rnn_mfcc_delete.argtypes = [vp]
    lib.rnn_mfcc_extract.restype = None
    lib.rnn_mfcc_extract.argtypes = [vp, c_float_p, C.c_int, c_float_p, C.c_int]
    lib.rnn_mfcc_extract_device.restype = None
    lib.rnn_mfcc_extract_device.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    lib.rnn_mfcc_tables.restype = None
    lib.rnn_mfcc_tables.argtypes = [vp, c_float_p, ip, ip, c_float_p, c_float_p, c_float_p]
    lib.rnn_cells_new_sharded.restype = vp
    lib.rnn_cells_new_sharded.argtypes = [P, C.c_int, C.c_int]
    lib.rnn_cells_delete.restype = None
    lib.rnn_cells_delete.argtypes = [vp]
    lib.rnn_cells_forget.restype = None
    lib.rnn_cells_forget.argtypes = [vp]
    lib.rnn_cells_rnnca_frame.restype = None
    lib.rnn_cells_rnnca_frame.argtypes = [vp, bp, bp, ip, C.c_int, ip, C.c_int, C.c_int, C.c_int]
    lib.rnn_cells_rnnca_run.restype = None
    lib.rnn_cells_rnnca_run.argtypes = [vp, bp, C.c_int, bp, ip, C.c_int, ip, C.c_int, C.c_int,
                                        C.c_int]
    lib.rnn_cells_get_hidden.restype = None
    lib.rnn_cells_get_hidden.argtypes = [vp, C.c_int, c_float_p]
    lib.rnn_batch_bptt_log.restype = None
    lib.rnn_batch_bptt_log.argtypes = [vp, C.POINTER(RnnBatchBpttLog)]
    lib.rnn_batch_pull.restype = None
    lib.rnn_batch_pull.argtypes = [vp]
    lib.rnn_b200_comm_unique_id.restype = C.c_int
    lib.rnn_b200_comm_unique_id.argtypes = [vp]
    lib.rnn_b200_comm_join.restype = C.c_int
    lib.rnn_b200_comm_join.argtypes = [vp, C.c_int, C.c_int]
    lib.rnn_b200_comm_leave.restype = None
    lib.rnn_b200_comm_leave.argtypes = []
    lib.rnn_b200_comm_size.restype = C.c_int
    lib.rnn_b200_comm_size.argtypes = []
    lib.rnn_batch_p2p_export.restype = C.c_int
    lib.rnn_batch_p2p_export.argtypes = [vp, vp]
    lib.rnn_batch_p2p_attach.restype = C.c_int
    lib.rnn_batch_p2p_attach.argtypes = [vp, vp, C.c_int, C.c_int]
    return lib


_lib = None


def load_library(path=None):
    """Load (once) and return the ctypes handle of librecur_b200.so."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            "%s is missing: build it with `make -C recur_b200/csrc` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no fallback implementation." % p)
    handle = C.CDLL(p, mode=os.RTLD_LOCAL)
    abi.declare_rnn_api(handle)
    _declare_b200(handle)
    if path is None:
        _lib = handle
    return handle


class _Lazy(object):
    """`from recur_b200 import lib` without loading at import time."""

    def __getattr__(self, name):
        return getattr(load_library(), name)


lib = _Lazy()
