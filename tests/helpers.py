"""Shared helpers for the tests: nets in either library, numpy views, text."""
import ctypes as C

import numpy as np

from recur_b200 import abi

STD_FLAGS = abi.RNN_NET_FLAG_STANDARD | abi.RNN_NET_FLAG_BPTT_ADAPTIVE_MIN_ERROR


def fptr(a):
    return a.ctypes.data_as(abi.c_float_p)


def u8ptr(a):
    return a.ctypes.data_as(abi.u8_p)


def arr(ptr, n):
    """numpy view (no copy) of n floats behind a ctypes float pointer."""
    return np.ctypeslib.as_array(ptr, shape=(n,))


def make_net(lib, input_size=7, hidden=13, output=5, depth=6, seed=3, lr=1e-3,
             momentum=0.95, flags=STD_FLAGS, activation=abi.RNN_RELU, noise=0.0,
             init=True):
    net = lib.rnn_new(input_size, hidden, output, flags, seed, None, depth, lr,
                      momentum, noise, activation)
    if init:
        lib.rnn_randomise_weights_auto(net)
    return net


def weights(net):
    n = net.contents
    return arr(n.ih_weights, n.ih_size), arr(n.ho_weights, n.ho_size)


def copy_weights(dst, src):
    """Install src's weights in dst (same shapes)."""
    dw, do = weights(dst)
    sw, so = weights(src)
    dw[:] = sw
    do[:] = so


def markov_text(n, n_symbols=42, seed=2):
    """Deterministic order-1 Markov chain over n_symbols."""
    rng = np.random.RandomState(seed)
    trans = rng.dirichlet(np.ones(n_symbols) * 0.3, size=n_symbols)
    cum = np.cumsum(trans, axis=1)
    u = rng.random_sample(n)
    out = np.empty(n, dtype=np.uint8)
    s = 0
    for i in range(n):
        s = int(np.searchsorted(cum[s], u[i]))
        if s >= n_symbols:
            s = n_symbols - 1
        out[i] = s
    return out


def softmax_error_host(ref_or_none, y, target):
    """badmaths.h softmax_best_guess + error[target] += 1 through numpy
    (used only to feed both libraries the same o_error)."""
    y = np.asarray(y, dtype=np.float32)
    e = np.exp(y - y.max())
    p = (e / e.sum()).astype(np.float32)
    err = -p
    err[target] += 1.0
    return err


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)
