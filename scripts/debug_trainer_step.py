"""Where does the rnnca trainer step part from the reference?  (debug aid)"""
import ctypes as C
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from recur_b200 import api, abi
from helpers import make_net, weights, arr, fptr, rel_err

lib = api.load_library()
ref = oracle.load_ref(strict=True)
n_tr, n_in = int(sys.argv[1]) if len(sys.argv) > 1 else 200, 35
flags = abi.RNN_NET_FLAG_STANDARD | abi.RNN_COND_USE_SCALE | abi.RNN_NET_FLAG_LOG_WEIGHT_SUM
shape = dict(input_size=n_in, hidden=51, output=3, depth=10, seed=11, lr=3e-3, momentum=0.5,
             flags=flags)
rs = np.random.RandomState(1)
r = make_net(ref, **shape)
a = make_net(lib, **shape)
rn = ref.rnn_new_training_set(r, n_tr)
an = lib.rnn_new_training_set(a, n_tr)
batch = lib.rnn_batch_new(an, n_tr)
for f in range(3):
    x = rs.random_sample((n_tr, n_in)).astype(np.float32)
    err = (rs.random_sample((n_tr, 4)).astype(np.float32) - 0.5) * 0.1
    err[:, 3] = 0
    ref.rnn_bptt_clear_deltas(r)
    for t in range(n_tr):
        c = rn[t].contents
        arr(c.real_inputs, n_in)[:] = x[t]
        ref.rnn_opinion(rn[t], None, 0.0)
        arr(c.bptt.contents.o_error, c.o_size)[:] = err[t, :c.o_size]
        ref.rnn_bptt_calc_deltas(rn[t], 1, None)
    lib.rnn_bptt_clear_deltas(a)
    lib.rnn_batch_set_inputs(batch, fptr(x))
    lib.rnn_batch_opinion(batch, 0.0)
    e = np.ascontiguousarray(err[:, :3])   # n x output_size
    lib.rnn_batch_set_errors(batch, fptr(e))
    lib.rnn_batch_calc_deltas(batch, 1)
    lib.rnn_batch_pull(batch)
    ba, br = a.contents.bptt.contents, r.contents.bptt.contents
    ca, cr = a.contents, r.contents
    print(f, "ih_delta", rel_err(arr(ba.ih_delta, ca.ih_size), arr(br.ih_delta, cr.ih_size)),
          "ho_delta", rel_err(arr(ba.ho_delta, ca.ho_size), arr(br.ho_delta, cr.ho_size)),
          "|ih_delta|", np.abs(arr(br.ih_delta, cr.ih_size)).max())
    depths = (C.c_int32 * n_tr)()
    lib.rnn_batch_bptt_depths(batch, depths)
    print("   depths", sorted(set(depths)))
    ref.rnn_apply_learning(r, abi.RNN_MOMENTUM_WEIGHTED, 0.5)
    lib.rnn_apply_learning(a, abi.RNN_MOMENTUM_WEIGHTED, 0.5)
    print("   after apply", [rel_err(p, q) for p, q in zip(weights(a), weights(r))])
    ref.rnn_condition_net(r)
    lib.rnn_condition_net(a)
    print("   after condition", [rel_err(p, q) for p, q in zip(weights(a), weights(r))])
