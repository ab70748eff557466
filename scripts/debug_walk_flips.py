"""Diagnostic (not a test): where the GPU walk and the reference's end at
different depths, print both sides' log values for those streams."""
import ctypes as C
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from recur_b200 import api
from helpers import (make_net, weights, arr, u8ptr, markov_text, transplant_training_set,
                     reference_walk_logs, executed_depth)

lib = api.load_library()
ref = oracle.load_ref(strict=True)
shape = dict(input_size=42, hidden=1023, output=42, depth=30)
n, warm, steps, lr = 64, 32, 8, 1e-6
boost = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
engine = int(sys.argv[2]) if len(sys.argv) > 2 else 2
text = markov_text(20000, 42, seed=2)
g = make_net(lib, seed=1, lr=lr, **shape)
r = make_net(ref, seed=1, lr=lr, **shape)
if boost != 1.0:
    for w in weights(g):
        w *= boost
gn = lib.rnn_new_training_set(g, n)
rn = ref.rnn_new_training_set(r, n)
lib.rnn_b200_set_engine(engine)
batch = lib.rnn_batch_new(gn, n)
lib.rnn_batch_text_upload(batch, u8ptr(text), len(text))
lib.rnn_batch_text_train(batch, 0, warm, 0, 0.95, 2000.0, None)
lib.rnn_batch_pull(batch)
transplant_training_set(gn, ref, rn, n)
tmp = tempfile.mkdtemp()
for s in range(steps):
    logs = reference_walk_logs(ref, rn, n, lambda: ref.ref_multi_tap_train(
        rn, n, u8ptr(text), len(text), warm + s, 1, 0, 0.95, 2000.0, None, None, None), tmp)
    lib.rnn_batch_text_train(batch, warm + s, 1, 0, 0.95, 2000.0, None)
    got = (api.RnnBatchBpttLog * n)()
    lib.rnn_batch_bptt_log(batch, got)
    lib.rnn_batch_pull(batch)
    I = g.contents.ih_size
    d = arr(g.contents.bptt.contents.ih_delta, I)
    e = arr(r.contents.bptt.contents.ih_delta, I)
    print("step %d: %s ih_delta rel err %.2e" % (s, lib.rnn_b200_last_walk_kernel().decode(),
                                                  np.abs(d - e).max() / np.abs(e).max()))
    worst = 0
    for j in range(n):
        l = logs[j]
        want = executed_depth(l["depth"], 30)
        es_ref = l["scaled_error"] / l["ih_scale"]
        rel = abs(got[j].error_sum - es_ref) / max(es_ref, 1e-30)
        worst = max(worst, rel) if got[j].n_steps == want else worst
        if got[j].n_steps != want:
            print("  stream %d: depth gpu %d ref %d | es gpu %.6g ref %.5g | thr gpu %.6g ref %.5g | "
                  "top gpu %.6g ref %.5g | cum gpu %.6g ref %.5g | mef gpu %.6g ref %.5g" % (
                      j, got[j].n_steps, want, got[j].error_sum, es_ref,
                      got[j].min_error_threshold, l["min_error_threshold"],
                      got[j].top_error_scaled, l["top_error_scaled"], got[j].cum_error,
                      l["cum_error"], got[j].min_error_factor, l["min_error_factor"]))
    print("  worst rel difference of the final error_sum among equal-depth streams: %.2e" % worst)
    transplant_training_set(gn, ref, rn, n)
