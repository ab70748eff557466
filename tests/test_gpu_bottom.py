"""nets with a layer below the recurrent one
(rnn_new_with_bottom_layer; recur-nn.c:88-103, 377-382, 395-401, 751-764),
through the per-net calls and through the array-of-nets calls, against the
reference — including its shared, ever-growing cumulative input error."""
import ctypes as C

import numpy as np
import pytest

from recur_b200 import abi
from helpers import weights, arr, fptr, rel_err, STD_FLAGS

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_bottom_net(lib, n_inputs=9, r_inputs=6, hidden=21, output=9, depth=5, seed=5, lr=0.02,
                    noise=0.0, boost=1.0):
    net = lib.rnn_new_with_bottom_layer(n_inputs, r_inputs, hidden, output, STD_FLAGS, seed, None,
                                        depth, lr, 0.9, noise, abi.RNN_RELU, 0)
    lib.rnn_randomise_weights_auto(net)
    if boost != 1.0:
        ih, ho = weights(net)
        ih *= boost
        ho *= boost
    return net


def bottom_weights(net):
    bl = net.contents.bottom_layer.contents
    return arr(bl.weights, bl.i_size * bl.o_size)


def one_hot_step_per_net(L, ref, nj, hot, target, accumulate, noise):
    """one_hot_opinion + net_error_bptt + calc_deltas for one stream
    (charmodel-helpers.h:16-33, charmodel-predict.c:18-27)."""
    c = nj.contents
    bl = c.bottom_layer.contents
    L.rnn_bptt_advance(nj)
    x = arr(bl.inputs, bl.input_size)
    x[:] = 0
    x[hot] = 1.0
    answer = L.rnn_opinion(nj, None, noise)
    ref.ref_softmax_best_guess(c.bptt.contents.o_error, answer, c.output_size)
    arr(c.bptt.contents.o_error, c.o_size)[target] += 1.0
    L.rnn_bptt_calc_deltas(nj, accumulate, None)


@pytest.mark.parametrize("noise,boost,lr", [(0.0, 1.0, 0.02), (0.05, 1.0, 0.02), (0.0, 3.0, 0.1)])
def test_bottom_layer_per_net_training(gpu_lib, ref, noise, boost, lr):
    lib = gpu_lib
    n, steps = 3, 8
    rs = np.random.RandomState(3)
    sym = rs.randint(1, 9, size=(steps + 1, n))
    res = []
    for L in (lib, ref):
        net = make_bottom_net(L, noise=noise, boost=boost, lr=lr)
        nets = L.rnn_new_training_set(net, n)
        for t in range(steps):
            for j in range(n):
                one_hot_step_per_net(L, ref, nets[j], sym[t, j], sym[t + 1, j], 1 if j else 0, noise)
            L.rnn_apply_learning(net, 0, 0.9)
        bl = net.contents.bottom_layer.contents
        res.append(dict(ih=weights(net)[0].copy(), ho=weights(net)[1].copy(),
                        bw=bottom_weights(net).copy(),
                        boe=arr(bl.o_error, bl.o_size).copy(),
                        hid=arr(nets[1].contents.hidden_layer, net.contents.h_size).copy(),
                        rng=(net.contents.rng.a, net.contents.rng.d)))
    for k in ("ih", "ho", "bw", "boe", "hid"):
        assert rel_err(res[0][k], res[1][k]) < TOL, k
    assert res[0]["rng"] == res[1]["rng"]


@pytest.mark.parametrize("n", [5, 64])
def test_bottom_layer_batch_training(gpu_lib, ref, n):
    """Dense feature rows through the bottom layer (the gstclassify / parrot
    shape) with the array-of-nets calls; the reference runs stream by stream."""
    lib = gpu_lib
    steps, F = 6, 12
    rs = np.random.RandomState(n)
    feats = rs.random_sample((steps, n, F)).astype(np.float32)
    tgt = rs.randint(0, 4, size=(steps, n))
    kw = dict(n_inputs=F, r_inputs=7, hidden=67, output=4, depth=4, seed=8, lr=0.01)
    r = make_bottom_net(ref, **kw)
    a = make_bottom_net(lib, **kw)
    rn = ref.rnn_new_training_set(r, n)
    an = lib.rnn_new_training_set(a, n)
    batch = lib.rnn_batch_new(an, n)
    for t in range(steps):
        ref.rnn_bptt_clear_deltas(r)
        for j in range(n):
            c = rn[j].contents
            out = ref.rnn_opinion(rn[j], fptr(feats[t, j]), 0.0)
            ref.ref_softmax_best_guess(c.bptt.contents.o_error, out, 4)
            arr(c.bptt.contents.o_error, c.o_size)[tgt[t, j]] += 1.0
            ref.rnn_bptt_calc_deltas(rn[j], 1, None)
            ref.rnn_bptt_advance(rn[j])
        ref.rnn_apply_learning(r, abi.RNN_MOMENTUM_NESTEROV, 0.9)
        lib.rnn_bptt_clear_deltas(a)
        lib.rnn_batch_set_inputs(batch, fptr(np.ascontiguousarray(feats[t])))
        lib.rnn_batch_opinion(batch, 0.0)
        lib.rnn_batch_softmax_error(batch, np.ascontiguousarray(tgt[t].astype(np.uint8)).ctypes.data_as(abi.u8_p),
                                    None, None)
        lib.rnn_batch_calc_deltas(batch, 1)
        lib.rnn_batch_advance(batch)
        lib.rnn_apply_learning(a, abi.RNN_MOMENTUM_NESTEROV, 0.9)
    for x, y in zip(weights(a), weights(r)):
        assert rel_err(x, y) < TOL
    assert rel_err(bottom_weights(a), bottom_weights(r)) < TOL
    bla, blr = a.contents.bottom_layer.contents, r.contents.bottom_layer.contents
    lib.rnn_b200_synchronize()
    assert rel_err(arr(bla.o_error, bla.o_size), arr(blr.o_error, blr.o_size)) < TOL
    lib.rnn_batch_delete(batch)
