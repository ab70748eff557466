"""TEST INFRASTRUCTURE ONLY.  Loaders for the two CPU oracles:

  ref   oracle/_ref/librecur_ref.so — the unmodified reference compiled in
        place (oracle/Makefile), driven through the same ctypes ABI as the
        product library;
  port  oracle/liboracle_rnn.so — this repo's plain-C restatement.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this package; the product (recur_b200/) never does.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FAST = os.path.join(HERE, "_ref", "librecur_ref.so")
REF_STRICT = os.path.join(HERE, "_ref", "librecur_ref_strict.so")
PORT = os.path.join(HERE, "liboracle_rnn.so")

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
u8_p = C.POINTER(C.c_uint8)


def build(ref=True, port=True):
    """Run oracle/Makefile (the reference part is a no-op where
    /root/reference is absent, e.g. on the GPU box)."""
    targets = []
    if port:
        targets.append("port")
    if ref:
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def have_ref():
    return os.path.exists(REF_FAST) and os.path.exists(REF_STRICT)


_cache = {}


def load_ref(strict=True):
    """The unmodified reference compiled in place, with the rnn_* API
    declared exactly as for the product library."""
    from recur_b200 import abi
    path = REF_STRICT if strict else REF_FAST
    if path in _cache:
        return _cache[path]
    lib = C.CDLL(path, mode=os.RTLD_LOCAL)
    abi.declare_rnn_api(lib)
    P = abi.RecurNN_p
    lib.ref_fast_expf.restype = C.c_float
    lib.ref_fast_expf.argtypes = [C.c_float]
    lib.ref_fast_sigmoid.restype = C.c_float
    lib.ref_fast_sigmoid.argtypes = [C.c_float]
    lib.ref_soft_clip.restype = C.c_float
    lib.ref_soft_clip.argtypes = [C.c_float, C.c_float]
    lib.ref_softmax.restype = None
    lib.ref_softmax.argtypes = [c_float_p, c_float_p, C.c_int]
    lib.ref_softmax_best_guess.restype = C.c_int
    lib.ref_softmax_best_guess.argtypes = [c_float_p, c_float_p, C.c_int]
    lib.ref_init_rand64.restype = None
    lib.ref_init_rand64.argtypes = [C.POINTER(abi.RandCtx), C.c_uint64]
    lib.ref_rand64.restype = C.c_uint64
    lib.ref_rand64.argtypes = [C.POINTER(abi.RandCtx)]
    lib.ref_rand_double.restype = C.c_double
    lib.ref_rand_double.argtypes = [C.POINTER(abi.RandCtx)]
    lib.ref_cheap_gaussian_noise.restype = C.c_float
    lib.ref_cheap_gaussian_noise.argtypes = [C.POINTER(abi.RandCtx)]
    lib.ref_one_hot_error.restype = C.c_float
    lib.ref_one_hot_error.argtypes = [P, C.c_int, C.c_int, c_int_p]
    lib.ref_multi_tap_train.restype = C.c_double
    lib.ref_multi_tap_train.argtypes = [abi.RecurNN_pp, C.c_int, u8_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_float, c_double_p, c_double_p, c_int_p]
    lib.ref_single_net_train.restype = C.c_double
    lib.ref_single_net_train.argtypes = [P, u8_p, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_float, C.c_uint,
                                         c_double_p, c_double_p, c_int_p]
    lib.ref_opinion_steps.restype = C.c_double
    lib.ref_opinion_steps.argtypes = [P, u8_p, C.c_int, C.c_int]
    lib.ref_rnnca_cells.restype = C.c_double
    lib.ref_rnnca_cells.argtypes = [C.POINTER(abi.RecurNN_p), C.c_int, C.c_int,
                                    C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int, C.c_int,
                                    C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int),
                                    C.c_int, C.c_int, C.c_int]
    lib.ref_multi_entropy_step.restype = None
    lib.ref_multi_entropy_step.argtypes = [c_float_p, C.c_int, C.c_int, C.c_int, c_double_p]
    lib.rnn_char_multi_cross_entropy.restype = None
    lib.rnn_char_multi_cross_entropy.argtypes = [P, u8_p, C.c_int, C.c_int, c_double_p, C.c_int]
    for name in ("ref_sizeof_RecurNN", "ref_sizeof_RecurNNBPTT",
                 "ref_sizeof_RecurExtraLayer"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = []
    _cache[path] = lib
    return lib


CHARMULTI_B200 = os.path.join(HERE, "_ref", "libcharmulti_b200.so")
FIXTURE_NET = os.path.join(HERE, "_ref", "fixtures", "multi-text-6c34c563i73-h99-o3650.net")


def load_charmulti_b200():
    """The reference's charmodel-multi-predict.c, unmodified, linked against
    librecur_b200.so (oracle/Makefile): its rnn_char_multi_cross_entropy calls
    this repo's rnn_opinion."""
    from recur_b200 import abi
    if CHARMULTI_B200 in _cache:
        return _cache[CHARMULTI_B200]
    lib = C.CDLL(CHARMULTI_B200, mode=os.RTLD_LOCAL)
    lib.rnn_char_multi_cross_entropy.restype = None
    lib.rnn_char_multi_cross_entropy.argtypes = [abi.RecurNN_p, u8_p, C.c_int, C.c_int,
                                                 c_double_p, C.c_int]
    _cache[CHARMULTI_B200] = lib
    return lib


class OracleDims(C.Structure):
    _fields_ = [("i_size", C.c_int), ("h_size", C.c_int), ("o_size", C.c_int),
                ("input_size", C.c_int), ("hidden_size", C.c_int),
                ("output_size", C.c_int)]


class OracleBpttResult(C.Structure):
    _fields_ = [("top_raw", C.c_float), ("top_scaled", C.c_float),
                ("err_sum", C.c_float), ("ih_scale", C.c_float),
                ("min_error_factor", C.c_float), ("cum_error", C.c_float),
                ("min_error_sum", C.c_float), ("n_steps", C.c_int),
                ("t_left", C.c_int)]


def load_port():
    """This repo's plain-C restatement (oracle/oracle_rnn.c)."""
    if PORT in _cache:
        return _cache[PORT]
    if not os.path.exists(PORT):
        build(ref=False, port=True)
    lib = C.CDLL(PORT, mode=os.RTLD_LOCAL)
    vp = C.c_void_p
    D = C.POINTER(OracleDims)
    lib.oracle_fast_expf.restype = C.c_float
    lib.oracle_fast_expf.argtypes = [C.c_float]
    lib.oracle_soft_clip.restype = C.c_float
    lib.oracle_soft_clip.argtypes = [C.c_float, C.c_float]
    lib.oracle_softmax_error.restype = C.c_float
    lib.oracle_softmax_error.argtypes = [c_float_p, C.c_int, C.c_int, c_float_p, c_int_p]
    lib.oracle_forward.restype = None
    lib.oracle_forward.argtypes = [D, c_float_p, c_float_p, c_float_p, c_float_p,
                                   c_float_p, C.c_int, c_float_p]
    lib.oracle_calc_deltas.restype = None
    lib.oracle_calc_deltas.argtypes = [D, c_float_p, c_float_p, c_float_p, C.c_int,
                                       C.c_int, c_float_p, c_float_p, C.c_float,
                                       C.c_float, C.c_int, C.c_int, c_float_p,
                                       c_float_p, c_float_p,
                                       C.POINTER(OracleBpttResult)]
    lib.oracle_apply_learning.restype = None
    lib.oracle_apply_learning.argtypes = [C.c_int, c_float_p, c_float_p, c_float_p,
                                          c_float_p, C.c_int, C.c_float, C.c_float,
                                          C.c_float]
    lib.oracle_rnnca_fill_inputs.restype = None
    lib.oracle_rnnca_fill_inputs.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int,
                                             c_int_p, C.c_int, c_int_p, C.c_int, C.c_int,
                                             C.c_int, c_float_p]
    lib.oracle_rnnca_unit_to_byte.restype = C.c_uint8
    lib.oracle_rnnca_unit_to_byte.argtypes = [C.c_float]
    lib.oracle_momentum_soft_start.restype = C.c_float
    lib.oracle_momentum_soft_start.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.oracle_set_new.restype = vp
    lib.oracle_set_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_int, C.c_int, c_float_p, c_float_p]
    lib.oracle_set_delete.restype = None
    lib.oracle_set_delete.argtypes = [vp]
    for name in ("oracle_set_wih", "oracle_set_who", "oracle_set_ih_delta",
                 "oracle_set_ho_delta", "oracle_set_hidden", "oracle_set_out",
                 "oracle_set_o_error", "oracle_set_mef"):
        getattr(lib, name).restype = c_float_p
        getattr(lib, name).argtypes = [vp]
    lib.oracle_set_char_step.restype = None
    lib.oracle_set_char_step.argtypes = [vp, u8_p, u8_p, C.c_float, c_double_p,
                                         c_double_p, c_int_p]
    lib.oracle_set_text_train.restype = C.c_double
    lib.oracle_set_text_train.argtypes = [vp, u8_p, C.c_int, C.c_int, C.c_int,
                                          C.c_float, C.c_float, c_double_p,
                                          c_double_p, c_int_p]
    _cache[PORT] = lib
    return lib
