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


def transplant_training_set(src_nets, dst_lib, dst_nets, n):
    """Copy everything a training step reads - weights, momentums, and every
    stream's ring, hidden layer, ring index, min_error_factor, generation -
    from the nets of one library (mirrors already pulled) into the nets of
    another with the same shapes.  This is the 'teacher forcing' of SURVEY.md
    §8c: both sides then take the next steps from bit-identical state, so a
    per-step comparison is not blurred by earlier rounding drift."""
    s0, d0 = src_nets[0].contents, dst_nets[0].contents
    sb, db = s0.bptt.contents, d0.bptt.contents
    for name, size in (("ih_weights", s0.ih_size), ("ho_weights", s0.ho_size)):
        arr(getattr(d0, name), size)[:] = arr(getattr(s0, name), size)
    for name, size in (("ih_momentum", s0.ih_size), ("ho_momentum", s0.ho_size)):
        arr(getattr(db, name), size)[:] = arr(getattr(sb, name), size)
    for j in range(n):
        s, d = src_nets[j].contents, dst_nets[j].contents
        sb, db = s.bptt.contents, d.bptt.contents
        assert sb.depth == db.depth and s.i_size == d.i_size
        arr(d.hidden_layer, d.h_size)[:] = arr(s.hidden_layer, s.h_size)
        arr(db.history, db.depth * d.i_size)[:] = arr(sb.history, sb.depth * s.i_size)
        # rnn_bptt_advance repoints input_layer / real_inputs into the ring
        # (recur-nn.c:696-704): step the destination onto the source's index
        db.index = (sb.index - 1) % sb.depth
        dst_lib.rnn_bptt_advance(dst_nets[j])
        assert db.index == sb.index
        db.min_error_factor = sb.min_error_factor
        db.ih_scale = sb.ih_scale
        db.learn_rate = sb.learn_rate
        d.generation = s.generation


def reference_walk_logs(ref, nets, n, run_step, tmpdir):
    """What the reference logs for ONE training step of every stream
    (recur-nn.c:415-421, 766-770): a list of dicts keyed by the log names.
    The log prints floats with %.5g: five significant digits."""
    import os
    paths = [os.path.join(str(tmpdir), "walk%d.log" % j) for j in range(n)]
    for j in range(n):
        ref.rnn_set_log_file(nets[j], paths[j].encode(), 0)
    run_step()
    out = []
    for j in range(n):
        ref.rnn_set_log_file(nets[j], None, 0)
        rec = {}
        for line in open(paths[j]).read().splitlines():
            key, _, val = line.partition(" ")
            if key != "generation":
                assert key not in rec, (j, key)     # exactly one walk was logged
            rec[key] = float(val)
        out.append(rec)
    return out


def executed_depth(logged_depth, depth):
    """recur-nn.c:416 logs depth - t, one less than the steps executed when the
    walk broke off early."""
    return min(int(logged_depth) + 1, depth)
