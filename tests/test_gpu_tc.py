"""The tensor-core engine (rb_tc.cu: tcgen05 3xTF32) against the plain-C
oracle and against the FMA engine, through the same C-ABI calls."""
import ctypes as C

import numpy as np
import pytest

from recur_b200 import abi, api
from helpers import make_net, weights, arr, fptr, u8ptr, markov_text, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def run_batch(lib, engine, shape, n, steps, text, lr, boost=1.0, pull=True):
    lib.rnn_b200_set_engine(engine)
    net = make_net(lib, seed=1, lr=lr, **shape)
    if boost != 1.0:
        ih, ho = weights(net)
        ih *= boost
        ho *= boost
    ih0, ho0 = [w.copy() for w in weights(net)]
    nets = lib.rnn_new_training_set(net, n)
    batch = lib.rnn_batch_new(nets, n)
    lib.rnn_batch_text_upload(batch, u8ptr(text), len(text))
    stats = api.RnnBatchCharStats()
    lib.rnn_batch_text_train(batch, 0, steps, 0, 0.95, 2000.0, C.byref(stats))
    lib.rnn_batch_pull(batch)
    H = net.contents.h_size
    res = dict(
        ih=weights(net)[0].copy(), ho=weights(net)[1].copy(),
        hidden=np.stack([arr(nets[j].contents.hidden_layer, H).copy() for j in range(n)]),
        mef=np.array([nets[j].contents.bptt.contents.min_error_factor for j in range(n)]),
        ih_scale=np.array([nets[j].contents.bptt.contents.ih_scale for j in range(n)]),
        ih_delta=arr(net.contents.bptt.contents.ih_delta, net.contents.ih_size).copy(),
        stats=(stats.error, stats.entropy, stats.correct, stats.count), ih0=ih0, ho0=ho0)
    lib.rnn_batch_delete(batch)
    lib.rnn_delete_training_set(nets, n, 0)
    lib.rnn_b200_set_engine(0)
    return res


@pytest.mark.parametrize("shape,n", [
    (dict(input_size=42, hidden=63, output=42, depth=6), 64),
    (dict(input_size=42, hidden=199, output=42, depth=30), 96),
    (dict(input_size=20, hidden=130, output=20, depth=9), 160),
])
def test_tensor_engine_matches_oracle_port(gpu_lib, port, shape, n):
    lib = gpu_lib
    steps = 5
    lr = 2e-4
    nsym = min(shape["input_size"], shape["output"])
    text = markov_text(3000, nsym, seed=2)
    got = run_batch(lib, 2, shape, n, steps, text, lr)
    s = port.oracle_set_new(shape["input_size"], shape["hidden"], shape["output"], n,
                            shape["depth"], lr, abi.RNN_RELU, 1, fptr(got["ih0"]), fptr(got["ho0"]))
    e, h, c = C.c_double(), C.c_double(), C.c_int()
    port.oracle_set_text_train(s, u8ptr(text), len(text), 0, steps, 0.95, 2000.0,
                               C.byref(e), C.byref(h), C.byref(c))
    H = got["hidden"].shape[1]
    assert rel_err(got["hidden"], arr(port.oracle_set_hidden(s), n * H).reshape(n, H)) < TOL
    assert rel_err(got["ih_delta"], arr(port.oracle_set_ih_delta(s), len(got["ih_delta"]))) < TOL
    assert rel_err(got["ih"], arr(port.oracle_set_wih(s), len(got["ih"]))) < TOL
    assert rel_err(got["ho"], arr(port.oracle_set_who(s), len(got["ho"]))) < TOL
    assert rel_err(got["mef"], arr(port.oracle_set_mef(s), n)) < TOL
    assert got["stats"][3] == n * steps
    assert got["stats"][2] == c.value
    assert abs(got["stats"][1] - h.value) < 1e-4 * abs(h.value)
    port.oracle_set_delete(s)


def test_tensor_and_fma_engines_agree_with_clipped_gradients(gpu_lib):
    """Doubled weights and a high learn rate clip ih_scale for some streams;
    the tensor engine re-splits those streams' error rows (k_finalize_rows)."""
    lib = gpu_lib
    shape = dict(input_size=12, hidden=75, output=12, depth=8)
    text = markov_text(2000, 12, seed=3)
    a = run_batch(lib, 1, shape, 64, 4, text, 0.01, boost=3.0)
    b = run_batch(lib, 2, shape, 64, 4, text, 0.01, boost=3.0)
    assert (a["ih_scale"] != 1).sum() >= 10   # the clip really happens
    for k in ("hidden", "ih", "ho", "ih_delta"):
        assert rel_err(b[k], a[k]) < TOL, k
    np.testing.assert_allclose(b["ih_scale"], a["ih_scale"], rtol=1e-3, atol=1e-5)


def test_forcing_the_tensor_engine_on_a_small_batch_aborts(gpu_lib):
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')\n"
            "from recur_b200 import api\nfrom helpers import *\n"
            "L = api.load_library()\nL.rnn_b200_set_engine(2)\n"
            "net = make_net(L)\nnets = L.rnn_new_training_set(net, 3)\n"
            "b = L.rnn_batch_new(nets, 3)\nL.rnn_batch_advance(b)\nL.rnn_batch_opinion(b, 0.0)\n"
            % (root, root))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "tensor engine was forced" in r.stderr
