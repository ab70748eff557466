"""ctypes view of the RecurNN C ABI (include/recur-nn.h).

The struct layouts restate reference recur-nn.h:158-227 (RecurNN, RecurNNBPTT,
RecurExtraLayer), recur-nn.h:230-258 (RecurInitialisationParameters),
recur-nn.h:262-265 (RecurErrorRange) and recur-rng.h:17-22 (rand_ctx).  The
same classes are used to drive this repo's librecur_b200.so and, in tests,
the reference compiled in place (oracle/_ref): both expose the identical
`rnn_*` symbol set, which is the point of the drop-in boundary.
"""
import ctypes as C

c_float_p = C.POINTER(C.c_float)
u8_p = C.POINTER(C.c_uint8)


class RandCtx(C.Structure):
    _fields_ = [("a", C.c_uint64), ("b", C.c_uint64),
                ("c", C.c_uint64), ("d", C.c_uint64)]


class RecurExtraLayer(C.Structure):
    _fields_ = [
        ("mem", c_float_p), ("weights", c_float_p), ("momentums", c_float_p),
        ("aux", c_float_p), ("delta", c_float_p), ("inputs", c_float_p),
        ("outputs", c_float_p), ("i_error", c_float_p), ("o_error", c_float_p),
        ("learn_rate_scale", C.c_float),
        ("input_size", C.c_int), ("output_size", C.c_int),
        ("i_size", C.c_int), ("o_size", C.c_int), ("overlap", C.c_int),
    ]


class RecurNNBPTT(C.Structure):
    _fields_ = [
        ("depth", C.c_int), ("index", C.c_int),
        ("i_error", c_float_p), ("h_error", c_float_p), ("o_error", c_float_p),
        ("ih_momentum", c_float_p), ("ho_momentum", c_float_p),
        ("history", c_float_p),
        ("ih_delta", c_float_p), ("ho_delta", c_float_p),
        ("ih_delta_tmp", c_float_p),
        ("ih_aux", c_float_p), ("ho_aux", c_float_p),
        ("mem", c_float_p),
        ("learn_rate", C.c_float), ("ih_scale", C.c_float),
        ("ho_scale", C.c_float), ("momentum", C.c_float),
        ("momentum_weight", C.c_float), ("min_error_factor", C.c_float),
    ]


class RecurNN(C.Structure):
    _fields_ = [
        ("i_size", C.c_int), ("h_size", C.c_int), ("o_size", C.c_int),
        ("input_size", C.c_int), ("hidden_size", C.c_int),
        ("output_size", C.c_int),
        ("ih_size", C.c_int), ("ho_size", C.c_int),
        ("flags", C.c_uint32),
        ("log", C.c_void_p),
        ("mem", c_float_p),
        ("input_layer", c_float_p), ("hidden_layer", c_float_p),
        ("output_layer", c_float_p),
        ("ih_weights", c_float_p), ("ho_weights", c_float_p),
        ("real_inputs", c_float_p),
        ("rng", RandCtx),
        ("bptt", C.POINTER(RecurNNBPTT)),
        ("bottom_layer", C.POINTER(RecurExtraLayer)),
        ("metadata", C.c_char_p),
        ("generation", C.c_uint32),
        ("presynaptic_noise", C.c_float),
        ("activation", C.c_int),
    ]


class RecurErrorRange(C.Structure):
    _fields_ = [("start", C.c_int), ("len", C.c_int)]


class RecurInitialisationParameters(C.Structure):
    _fields_ = [
        ("method", C.c_int), ("submethod", C.c_int),
        ("bias_uses_submethod", C.c_int), ("inputs_use_submethod", C.c_int),
        ("fan_in_sum", C.c_float), ("fan_in_step", C.c_float),
        ("fan_in_min", C.c_float), ("fan_in_ratio", C.c_float),
        ("flat_variance", C.c_float), ("flat_shape", C.c_int),
        ("flat_perforation", C.c_double),
        ("run_input_probability", C.c_float), ("run_input_magnitude", C.c_float),
        ("run_gain", C.c_float), ("run_len_mean", C.c_float),
        ("run_len_stddev", C.c_float),
        ("run_n", C.c_int), ("run_loop", C.c_int),
        ("run_crossing_paths", C.c_int), ("run_inputs_miss", C.c_int),
        ("run_input_at_start", C.c_int),
    ]


RecurNN_p = C.POINTER(RecurNN)
RecurNN_pp = C.POINTER(RecurNN_p)

# flags (recur-nn.h:78-103)
RNN_NET_FLAG_OWN_BPTT = 1
RNN_NET_FLAG_OWN_WEIGHTS = 2
RNN_NET_FLAG_LOG_APPEND = 8
RNN_NET_FLAG_LOG_HIDDEN_SUM = 16
RNN_NET_FLAG_LOG_WEIGHT_SUM = 32
RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR = 64
RNN_NET_FLAG_NO_MOMENTUMS = 128
RNN_NET_FLAG_NO_DELTAS = 256
RNN_NET_FLAG_BOTTOM_LAYER = 1024
RNN_NET_FLAG_AUX_ARRAYS = 2048
RNN_COND_USE_SCALE = 1 << 16
RNN_COND_USE_ZERO = 1 << 18
RNN_COND_USE_LAWN_MOWER = 1 << 19
RNN_COND_USE_TALL_POPPY = 1 << 20
RNN_COND_USE_RAND = 1 << 22
RNN_NET_FLAG_STANDARD = (RNN_NET_FLAG_OWN_BPTT | RNN_NET_FLAG_OWN_WEIGHTS |
                         RNN_COND_USE_ZERO | RNN_NET_FLAG_LOG_HIDDEN_SUM)

# learning methods (recur-nn.h:109-119)
RNN_MOMENTUM_WEIGHTED = 0
RNN_MOMENTUM_NESTEROV = 1
RNN_MOMENTUM_SIMPLIFIED_NESTEROV = 2
RNN_MOMENTUM_CLASSICAL = 3
RNN_ADAGRAD = 4
RNN_ADADELTA = 5
RNN_RPROP = 6

# activations (recur-nn.h:130-140)
RNN_RELU = 1
RNN_RESQRT = 2
RNN_RECLIP20 = 5

# init methods (recur-nn.h:121-128)
RNN_INIT_ZERO, RNN_INIT_FLAT, RNN_INIT_FAN_IN, RNN_INIT_RUNS = 0, 1, 2, 3

RECUR_RNG_SUBSEED = (1 << 64) - 2


def declare_rnn_api(lib):
    """Attach argtypes/restypes for the rnn_* functions of recur-nn.h:269-334
    to a loaded shared library (ours or the compiled reference)."""
    f = lib.rnn_new
    f.restype = RecurNN_p
    f.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint32, C.c_uint64,
                  C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
    f = lib.rnn_clone
    f.restype = RecurNN_p
    f.argtypes = [RecurNN_p, C.c_uint32, C.c_uint64, C.c_char_p]
    f = lib.rnn_new_extra_layer
    f.restype = C.POINTER(RecurExtraLayer)
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32]
    f = lib.rnn_new_with_bottom_layer
    f.restype = RecurNN_p
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint64,
                  C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_float,
                  C.c_int, C.c_int]
    lib.rnn_set_log_file.restype = None
    lib.rnn_set_log_file.argtypes = [RecurNN_p, C.c_char_p, C.c_int]
    lib.rnn_randomise_weights_clever.restype = None
    lib.rnn_randomise_weights_clever.argtypes = [
        RecurNN_p, C.POINTER(RecurInitialisationParameters)]
    lib.rnn_randomise_weights_simple.restype = None
    lib.rnn_randomise_weights_simple.argtypes = [RecurNN_p, C.c_int]
    lib.rnn_randomise_weights_auto.restype = None
    lib.rnn_randomise_weights_auto.argtypes = [RecurNN_p]
    lib.rnn_init_default_weight_parameters.restype = None
    lib.rnn_init_default_weight_parameters.argtypes = [
        RecurNN_p, C.POINTER(RecurInitialisationParameters)]
    lib.rnn_scale_initial_weights.restype = None
    lib.rnn_scale_initial_weights.argtypes = [RecurNN_p, C.c_float]
    lib.rnn_print_net_stats.restype = None
    lib.rnn_print_net_stats.argtypes = [RecurNN_p]
    lib.rnn_delete_net.restype = None
    lib.rnn_delete_net.argtypes = [RecurNN_p]
    lib.rnn_new_training_set.restype = RecurNN_pp
    lib.rnn_new_training_set.argtypes = [RecurNN_p, C.c_int]
    lib.rnn_delete_training_set.restype = None
    lib.rnn_delete_training_set.argtypes = [RecurNN_pp, C.c_int, C.c_int]
    lib.rnn_opinion.restype = c_float_p
    lib.rnn_opinion.argtypes = [RecurNN_p, c_float_p, C.c_float]
    lib.rnn_multi_pgm_dump.restype = None
    lib.rnn_multi_pgm_dump.argtypes = [RecurNN_p, C.c_char_p, C.c_char_p]
    lib.rnn_load_net.restype = RecurNN_p
    lib.rnn_load_net.argtypes = [C.c_char_p]
    lib.rnn_save_net.restype = C.c_int
    lib.rnn_save_net.argtypes = [RecurNN_p, C.c_char_p, C.c_int]
    lib.rnn_bptt_clear_deltas.restype = None
    lib.rnn_bptt_clear_deltas.argtypes = [RecurNN_p]
    lib.rnn_bptt_advance.restype = None
    lib.rnn_bptt_advance.argtypes = [RecurNN_p]
    lib.rnn_bptt_calculate.restype = None
    lib.rnn_bptt_calculate.argtypes = [RecurNN_p, C.c_uint]
    lib.rnn_apply_learning.restype = None
    lib.rnn_apply_learning.argtypes = [RecurNN_p, C.c_int, C.c_float]
    lib.rnn_calculate_momentum_soft_start.restype = C.c_float
    lib.rnn_calculate_momentum_soft_start.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.rnn_bptt_calc_deltas.restype = None
    lib.rnn_bptt_calc_deltas.argtypes = [RecurNN_p, C.c_int,
                                         C.POINTER(RecurErrorRange)]
    lib.rnn_condition_net.restype = None
    lib.rnn_condition_net.argtypes = [RecurNN_p]
    lib.rnn_log_net.restype = None
    lib.rnn_log_net.argtypes = [RecurNN_p]
    lib.rnn_forget_history.restype = None
    lib.rnn_forget_history.argtypes = [RecurNN_p, C.c_int]
    lib.rnn_perforate_weights.restype = None
    lib.rnn_perforate_weights.argtypes = [RecurNN_p, C.c_float]
    lib.rnn_weight_noise.restype = None
    lib.rnn_weight_noise.argtypes = [RecurNN_p, C.c_float]
    lib.rnn_set_momentum_values.restype = None
    lib.rnn_set_momentum_values.argtypes = [RecurNN_p, C.c_float]
    lib.rnn_set_aux_values.restype = None
    lib.rnn_set_aux_values.argtypes = [RecurNN_p, C.c_float]
    lib.rnn_zap_non_diagonals.restype = None
    lib.rnn_zap_non_diagonals.argtypes = [RecurNN_p, C.c_int, C.c_int, C.c_int]
    lib.rnn_clear_diagonal_only_section.restype = None
    lib.rnn_clear_diagonal_only_section.argtypes = [RecurNN_p, C.c_uint, C.c_uint]
    return lib


RNN_API_SYMBOLS = [
    "rnn_new", "rnn_clone", "rnn_new_extra_layer", "rnn_new_with_bottom_layer",
    "rnn_set_log_file", "rnn_randomise_weights_clever",
    "rnn_randomise_weights_simple", "rnn_randomise_weights_auto",
    "rnn_init_default_weight_parameters", "rnn_scale_initial_weights",
    "rnn_print_net_stats", "rnn_delete_net", "rnn_new_training_set",
    "rnn_delete_training_set", "rnn_opinion", "rnn_multi_pgm_dump",
    "rnn_load_net", "rnn_save_net", "rnn_bptt_clear_deltas", "rnn_bptt_advance",
    "rnn_bptt_calculate", "rnn_apply_learning",
    "rnn_calculate_momentum_soft_start", "rnn_bptt_calc_deltas",
    "rnn_condition_net", "rnn_log_net", "rnn_forget_history",
    "rnn_perforate_weights", "rnn_weight_noise", "rnn_set_momentum_values",
    "rnn_set_aux_values", "rnn_zap_non_diagonals",
    "rnn_clear_diagonal_only_section",
]
