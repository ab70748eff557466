"""Short drivers for ncu captures of the single-stream and resident-weights
kernels (profiles/): `python scripts/ncu_small_kernels.py single` runs 60
characters of config 1 through the per-net API, `... resident` runs 12 chunks of
config 3 (256 channels, H199)."""
import sys, ctypes as C
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import oracle
from recur_b200 import api, abi
from helpers import make_net, fptr, arr, markov_text, STD_FLAGS
lib = api.load_library()
ref = oracle.load_ref(strict=False)
mode = sys.argv[1]
if mode == "single":
    text = markov_text(2000, 42, seed=6)
    a = make_net(lib, input_size=42, hidden=199, output=42, depth=30, seed=1, lr=1e-3)
    c = a.contents
    o_err = c.bptt.contents.o_error
    for i in range(60):
        c.bptt.contents.momentum = lib.rnn_calculate_momentum_soft_start(c.generation, 0.95, 2000.0)
        lib.rnn_bptt_advance(a)
        x = arr(c.real_inputs, c.input_size)
        x[:] = 0
        x[text[i]] = 1.0
        ans = lib.rnn_opinion(a, None, 0.0)
        ref.ref_softmax_best_guess(o_err, ans, c.output_size)
        o_err[text[i + 1]] += 1.0
        lib.rnn_bptt_calculate(a, 1)
else:
    B, F, classes = 256, 32, 4
    rs = np.random.RandomState(3)
    a = make_net(lib, input_size=F, hidden=199, output=classes, depth=30, seed=11, lr=1e-6, flags=STD_FLAGS)
    an = lib.rnn_new_training_set(a, B)
    batch = lib.rnn_batch_new(an, B)
    outs = np.zeros((B, classes), dtype=np.float32)
    for t in range(12):
        feats = np.log1p(rs.random_sample((B, F)) * 400).astype(np.float32)
        targets = rs.randint(0, classes, size=B)
        lib.rnn_batch_set_inputs(batch, fptr(feats))
        lib.rnn_batch_opinion(batch, 0.0)
        lib.rnn_batch_get_outputs(batch, fptr(outs))
        e = np.exp(outs - outs.max(axis=1, keepdims=True)); p = e / e.sum(axis=1, keepdims=True)
        err = -p; err[np.arange(B), targets] += 1.0
        lib.rnn_batch_set_errors(batch, fptr(np.ascontiguousarray(err.astype(np.float32))))
        lib.rnn_batch_calc_deltas(batch, 0)
        lib.rnn_batch_advance(batch)
        lib.rnn_apply_learning(a, abi.RNN_MOMENTUM_NESTEROV, 0.9)
        lib.rnn_condition_net(a)
