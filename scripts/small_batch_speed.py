"""text-predict multi-tap at small batch sizes (below the tensor engine): chars/s"""
import sys, time, ctypes as C
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from recur_b200 import api, abi
from helpers import make_net, u8ptr, markov_text
lib = api.load_library()
text = markov_text(200000, 42, seed=6)
for n in (8, 32, 48):
    net = make_net(lib, input_size=42, hidden=199, output=42, depth=30, seed=1, lr=1e-3 / n)
    nets = lib.rnn_new_training_set(net, n)
    batch = lib.rnn_batch_new(nets, n)
    lib.rnn_batch_text_upload(batch, u8ptr(text), len(text))
    pos = lib.rnn_batch_text_train(batch, 0, 50, 0, 0.95, 2000.0, None)
    lib.rnn_b200_synchronize()
    l0 = lib.rnn_b200_kernel_launches()
    t0 = time.perf_counter()
    pos = lib.rnn_batch_text_train(batch, pos, 300, 0, 0.95, 2000.0, None)
    lib.rnn_b200_synchronize()
    t1 = time.perf_counter()
    print("n=%d: %.0f chars/s, %.1f us/step, %.1f launches/step" % (n, 300 * n / (t1 - t0), (t1 - t0) / 300 * 1e6, (lib.rnn_b200_kernel_launches() - l0) / 300))
