"""The north star's long-run criterion: the cross-entropy trajectory of
text-predict training stays within 1% of the reference's after 1M characters.

The reference (compiled in place from /root/reference into oracle/_ref, fast
build: the flags the project itself ships with) trains 64 synchronic streams
on a Markov text with rnn_char_epoch's loop (charmodel-predict.c:288-311);
ours runs the same job through rnn_batch_text_train on the tensor engine.
Individual weights drift apart over a million characters (fp32 summation
order), the learning curve must not."""
import ctypes as C

import numpy as np
import pytest

from helpers import make_net, markov_text, u8ptr
from recur_b200 import api

pytestmark = pytest.mark.gpu

N_STREAMS = 64
WINDOW = 2048
N_WINDOWS = 8            # 64 * 2048 * 8 = 1,048,576 characters
SHAPE = dict(input_size=42, hidden=63, output=42, depth=10, seed=3, lr=3e-4)


def test_entropy_trajectory_within_one_percent_after_1m_chars(gpu_lib, ref_fast):
    lib, ref = gpu_lib, ref_fast
    n = N_STREAMS
    text = markov_text(300000, 42, seed=5)
    r = make_net(ref, **SHAPE)
    g = make_net(lib, **SHAPE)
    rn = ref.rnn_new_training_set(r, n)
    gn = lib.rnn_new_training_set(g, n)
    batch = lib.rnn_batch_new(gn, n)
    lib.rnn_batch_text_upload(batch, u8ptr(text), len(text))
    pos_r = pos_g = 0
    want, got = [], []
    for w in range(N_WINDOWS):
        e, h, c = C.c_double(), C.c_double(), C.c_int()
        ref.ref_multi_tap_train(rn, n, u8ptr(text), len(text), pos_r, WINDOW, 0, 0.95, 2000.0,
                                C.byref(e), C.byref(h), C.byref(c))
        pos_r += WINDOW
        want.append((-h.value / (n * WINDOW), c.value / (n * WINDOW)))
        stats = api.RnnBatchCharStats()
        pos_g = lib.rnn_batch_text_train(batch, pos_g, WINDOW, 0, 0.95, 2000.0, C.byref(stats))
        assert stats.count == n * WINDOW
        got.append((-stats.entropy / stats.count, stats.correct / stats.count))
    want, got = np.array(want), np.array(got)
    # the reference really learns on this job: the curve is not flat
    assert want[-1, 0] < want[0, 0] - 1.0
    rel = np.abs(got[:, 0] - want[:, 0]) / want[:, 0]
    assert rel.max() < 0.01, (got, want)
    assert abs(got[-1, 0] - want[-1, 0]) < 0.01 * want[-1, 0]
    # accuracy follows too (absolute, it is a small number)
    assert np.abs(got[:, 1] - want[:, 1]).max() < 0.01
    lib.rnn_batch_delete(batch)
    lib.rnn_delete_training_set(gn, n, 0)
    ref.rnn_delete_training_set(rn, n, 0)
