"""Pins the plain-C oracle port (oracle/oracle_rnn.c) to the reference:
against the compiled reference itself where it is available, and against
the golden vectors generated from it (tests/golden/make_golden.py)."""
import ctypes as C
import os

import numpy as np

import oracle
from recur_b200 import abi
from helpers import fptr, u8ptr, arr, weights, make_net, markov_text

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "recur_golden.npz")


def test_fast_expf_against_golden(port):
    g = np.load(GOLDEN)
    got = np.array([port.oracle_fast_expf(float(x)) for x in g["expf_x"]], dtype=np.float32)
    # IEEE on both sides: identical
    assert np.array_equal(got, g["expf_y"])


def test_soft_clip_against_golden(port):
    g = np.load(GOLDEN)
    got = np.array([port.oracle_soft_clip(float(a), float(b)) for a, b in g["softclip_in"]],
                   dtype=np.float32)
    assert np.array_equal(got, g["softclip_out"])


def test_softmax_error_against_golden(port):
    g = np.load(GOLDEN)
    for k in range(int(g["softmax_n"])):
        y = g["softmax_y_%d" % k]
        err = np.zeros_like(y)
        w = C.c_int(-1)
        port.oracle_softmax_error(fptr(y), len(y), -1, fptr(err), C.byref(w))
        np.testing.assert_allclose(err, g["softmax_e_%d" % k], rtol=1e-6, atol=1e-30)
        assert w.value == int(g["softmax_w_%d" % k])


def replay_trace(port, g, prefix, lr, boost=1.0):
    """Run the port over the golden trace's text from the golden initial
    weights and compare every recorded quantity."""
    ih0, ho0 = g[prefix + "ih_weights0"].copy(), g[prefix + "ho_weights0"].copy()
    n_steps, n = g[prefix + "hidden"].shape[:2]
    s = port.oracle_set_new(7, 13, 7, n, 6, lr, abi.RNN_RELU, 1, fptr(ih0), fptr(ho0))
    text = g[prefix + "text"]
    spacing = (len(text) - 1) // n
    d_ih, d_ho = len(ih0), len(ho0)
    worst = 0.0
    for i in range(n_steps):
        cur = np.array([text[(i + j * spacing) % (len(text) - 1)] for j in range(n)], dtype=np.uint8)
        nxt = np.array([text[(i + j * spacing) % (len(text) - 1) + 1] for j in range(n)], dtype=np.uint8)
        port.oracle_set_char_step(s, u8ptr(cur), u8ptr(nxt), 0.9, None, None, None)
        hid = arr(port.oracle_set_hidden(s), n * 16).reshape(n, 16)
        out = arr(port.oracle_set_out(s), n * 8).reshape(n, 8)
        oe = arr(port.oracle_set_o_error(s), n * 8).reshape(n, 8)
        for name, got, want in (("hidden", hid, g[prefix + "hidden"][i]),
                                ("output", out, g[prefix + "output"][i]),
                                ("o_error", oe, g[prefix + "o_error"][i]),
                                ("mef", arr(port.oracle_set_mef(s), n), g[prefix + "mef"][i]),
                                ("ih_delta", arr(port.oracle_set_ih_delta(s), d_ih), g[prefix + "ih_delta"][i]),
                                ("ho_delta", arr(port.oracle_set_ho_delta(s), d_ho), g[prefix + "ho_delta"][i]),
                                ("ih_weights", arr(port.oracle_set_wih(s), d_ih), g[prefix + "ih_weights"][i]),
                                ("ho_weights", arr(port.oracle_set_who(s), d_ho), g[prefix + "ho_weights"][i])):
            scale = max(np.abs(want).max(), 1e-30)
            err = np.abs(got - want).max() / scale
            worst = max(worst, err)
            assert err < 2e-5, (prefix, i, name, err)
    port.oracle_set_delete(s)
    return worst


def test_port_replays_golden_trace(port):
    g = np.load(GOLDEN)
    replay_trace(port, g, "trace_", 0.02)


def test_port_replays_golden_trace_with_gradient_clipping(port):
    """The 'hot' trace clips ih_scale on 13 of 120 calls (down to 0.35)."""
    g = np.load(GOLDEN)
    assert (g["hot_ih_scale"] != 1).sum() >= 10
    replay_trace(port, g, "hot_", 0.1)


def test_port_against_live_reference_larger_net(port, ref):
    """Free-running comparison at text-predict's default shape (H199, D30)
    for a few steps, 4 streams: the chaotic divergence has no time to grow."""
    from helpers import markov_text
    n, steps = 4, 6
    net = make_net(ref, input_size=42, hidden=199, output=42, depth=30, seed=1, lr=1e-3)
    nets = ref.rnn_new_training_set(net, n)
    ih, ho = weights(net)
    s = port.oracle_set_new(42, 199, 42, n, 30, 1e-3, abi.RNN_RELU, 1, fptr(ih.copy()), fptr(ho.copy()))
    text = markov_text(3000, 42, seed=2)
    e1, h1, c1 = C.c_double(), C.c_double(), C.c_int()
    e2, h2, c2 = C.c_double(), C.c_double(), C.c_int()
    ref.ref_multi_tap_train(nets, n, u8ptr(text), len(text), 0, steps, 0, 0.95, 2000.0,
                            C.byref(e1), C.byref(h1), C.byref(c1))
    port.oracle_set_text_train(s, u8ptr(text), len(text), 0, steps, 0.95, 2000.0,
                               C.byref(e2), C.byref(h2), C.byref(c2))
    ih2 = arr(port.oracle_set_wih(s), len(ih))
    ho2 = arr(port.oracle_set_who(s), len(ho))
    assert np.abs(ih2 - ih).max() / np.abs(ih).max() < 1e-5
    assert np.abs(ho2 - ho).max() / np.abs(ho).max() < 1e-5
    assert abs(e1.value - e2.value) < 1e-4 * abs(e1.value)
    assert abs(h1.value - h2.value) < 1e-4 * abs(h1.value)
    assert c1.value == c2.value
    port.oracle_set_delete(s)


def test_reference_builds_agree_within_noise_floor(ref, ref_fast):
    """-Ofast/-ffast-math/-DVECTOR against the IEEE build: the reference's own
    wobble, which bounds how tight any parity tolerance can be."""
    from helpers import markov_text
    text = markov_text(2000, 42, seed=2)
    res = []
    for L in (ref, ref_fast):
        net = make_net(L, input_size=42, hidden=199, output=42, depth=30, seed=1, lr=1e-3)
        nets = L.rnn_new_training_set(net, 3)
        L.ref_multi_tap_train(nets, 3, u8ptr(text), len(text), 0, 5, 0, 0.95, 2000.0,
                              None, None, None)
        res.append([w.copy() for w in weights(net)])
    for a, b in zip(*res):
        assert np.abs(a - b).max() / np.abs(a).max() < 1e-5


def test_rnnca_gather_restatement_against_numpy(port):
    """f4: the restated fill_net_inputs / get_offset_point (gstrnnca.c:644-691,
    oracle_rnn.c) against an independent numpy statement of the same rule -
    clamp at the edges or wrap once, 1/255, position, radial term.  (The
    GStreamer element itself cannot be compiled; this pins the restatement to
    a second reading, not to the reference binary.)"""
    import ctypes as C
    rs = np.random.RandomState(4)
    W, H = 13, 9
    frame = rs.randint(0, 256, size=3 * W * H).astype(np.uint8)
    Y, Cb, Cr = frame.reshape(3, H, W)
    off_y = rs.randint(-2, 3, size=(17, 2)).astype(np.int32)
    off_c = rs.randint(-2, 3, size=(8, 2)).astype(np.int32)
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    for edges, len_pos in ((1, 2), (0, 3), (0, 2)):
        for cx, cy in ((0, 0), (W - 1, H - 1), (5, 4), (0, H - 1), (W - 1, 0)):
            got = np.zeros(17 + 16 + len_pos, dtype=np.float32)
            port.oracle_rnnca_fill_inputs(frame.ctypes.data_as(u8p), W, H, cx, cy,
                                          off_y.ctypes.data_as(ip), 17, off_c.ctypes.data_as(ip), 8,
                                          len_pos, edges, fptr(got))

            def at(plane, dx, dy):
                x, y = cx + dx, cy + dy
                if edges:
                    x, y = min(max(x, 0), W - 1), min(max(y, 0), H - 1)
                else:
                    x, y = x % W, y % H     # offsets are smaller than the frame: one wrap
                return np.float32(plane[y, x]) * np.float32(1.0 / 255.0)
            want = [at(Y, dx, dy) for dx, dy in off_y]
            for dx, dy in off_c:
                want += [at(Cb, dx, dy), at(Cr, dx, dy)]
            xx, yy = np.float32(cx) / np.float32(W), np.float32(cy) / np.float32(H)
            want += [xx, yy]
            if len_pos == 3:
                want.append(np.float32(0.5 - ((float(yy) - 0.5) ** 2 + (float(xx) - 0.5) ** 2)))
            np.testing.assert_allclose(got, np.array(want, dtype=np.float32), rtol=1e-6, atol=1e-7)
    assert port.oracle_rnnca_unit_to_byte(0.0) == 0
    assert port.oracle_rnnca_unit_to_byte(0.5) == 127
    assert port.oracle_rnnca_unit_to_byte(0.99999) == 255


def test_transplanted_training_state_continues_bit_for_bit(ref):
    """helpers.transplant_training_set (used by the GPU parity tests to start
    the reference from a state trained on the GPU) moves everything a
    training step reads: a reference training set continued in place and its
    transplanted copy take bit-identical steps."""
    from helpers import transplant_training_set
    n, shape = 3, dict(input_size=12, hidden=31, output=12, depth=7, seed=4, lr=3e-3)
    text = markov_text(900, 12, seed=4)
    a = make_net(ref, **shape)
    b = make_net(ref, **shape)
    for w in weights(b):
        w[:] = 0
    an = ref.rnn_new_training_set(a, n)
    bn = ref.rnn_new_training_set(b, n)
    ref.ref_multi_tap_train(an, n, u8ptr(text), len(text), 0, 11, 0, 0.95, 20.0, None, None, None)
    transplant_training_set(an, ref, bn, n)
    for nets in (an, bn):
        ref.ref_multi_tap_train(nets, n, u8ptr(text), len(text), 11, 4, 0, 0.95, 20.0,
                                None, None, None)
    for x, y in zip(weights(a), weights(b)):
        assert np.array_equal(x, y)
    for j in range(n):
        ca, cb = an[j].contents, bn[j].contents
        assert np.array_equal(arr(ca.hidden_layer, ca.h_size), arr(cb.hidden_layer, cb.h_size))
        assert ca.bptt.contents.min_error_factor == cb.bptt.contents.min_error_factor
