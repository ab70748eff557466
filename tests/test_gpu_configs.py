"""The other BASELINE.json configurations as parity cases: the stream loops of
gstclassify (config 3), the multi-head charmodel forward (config 4) and the
rnnca per-cell forward (config 5), replayed over the reference API on the CPU
(oracle/_ref) and over this library's array-of-nets calls on the GPU.  The
elements themselves need GStreamer and cannot be built, so
the loops are restated here in the order the reference runs them."""
import ctypes as C
import os

import numpy as np
import pytest

from recur_b200 import abi
from helpers import make_net, weights, arr, fptr, u8ptr, rel_err, copy_weights, STD_FLAGS

pytestmark = pytest.mark.gpu
TOL = 1e-4


def grouped_softmax_error(out, target):
    """train_channel's error for one class group (gstclassify.c:2104-2124):
    error = onehot(target) - softmax(outputs)."""
    e = np.exp(out - out.max())
    p = e / e.sum()
    err = -p
    err[target] += 1.0
    return err.astype(np.float32)


@pytest.mark.parametrize("n_channels,depth,chunks", [(10, 10, 6), (96, 10, 6), (256, 30, 34)])
def test_config3_classify_training_loop(gpu_lib, ref, n_channels, depth, chunks):
    """gstclassify.c:2201-2239: clear deltas; per channel forward on dense
    features, grouped softmax error, rnn_bptt_calc_deltas(net, 1, NULL),
    advance; then Nesterov update and rnn_condition_net.  The last case is
    BASELINE.json configs[2] at its own size: 256 channels, BPTT depth 30, run
    until the ring is full and the walks are as deep as they get."""
    lib = gpu_lib
    F, Hn, classes = 32, 199, 4
    rs = np.random.RandomState(3)
    feats = np.log1p(rs.random_sample((chunks, n_channels, F)) * 400).astype(np.float32)
    targets = rs.randint(0, classes, size=(chunks, n_channels))
    flags = STD_FLAGS
    r = make_net(ref, input_size=F, hidden=Hn, output=classes, depth=depth, seed=11, lr=1e-5,
                 flags=flags)
    a = make_net(lib, input_size=F, hidden=Hn, output=classes, depth=depth, seed=11, lr=1e-5,
                 flags=flags)
    rn = ref.rnn_new_training_set(r, n_channels)
    an = lib.rnn_new_training_set(a, n_channels)
    batch = lib.rnn_batch_new(an, n_channels)
    for t in range(chunks):
        # reference, channel by channel
        ref.rnn_bptt_clear_deltas(r)
        for j in range(n_channels):
            c = rn[j].contents
            out = ref.rnn_opinion(rn[j], fptr(feats[t, j]), 0.0)
            err = arr(c.bptt.contents.o_error, c.o_size)
            err[:classes] = grouped_softmax_error(arr(out, classes), targets[t, j])
            ref.rnn_bptt_calc_deltas(rn[j], 1, None)
            ref.rnn_bptt_advance(rn[j])
        ref.rnn_apply_learning(r, abi.RNN_MOMENTUM_NESTEROV, 0.9)
        ref.rnn_condition_net(r)
        # this library, all channels at once
        lib.rnn_batch_set_inputs(batch, fptr(np.ascontiguousarray(feats[t])))
        lib.rnn_batch_opinion(batch, 0.0)
        outs = np.zeros((n_channels, classes), dtype=np.float32)
        lib.rnn_batch_get_outputs(batch, fptr(outs))
        errs = np.stack([grouped_softmax_error(outs[j], targets[t, j]) for j in range(n_channels)])
        lib.rnn_batch_set_errors(batch, fptr(np.ascontiguousarray(errs)))
        lib.rnn_batch_calc_deltas(batch, 0)
        lib.rnn_batch_advance(batch)
        lib.rnn_apply_learning(a, abi.RNN_MOMENTUM_NESTEROV, 0.9)
        lib.rnn_condition_net(a)
        ref_outs = np.stack([arr(rn[j].contents.output_layer, classes).copy()
                             for j in range(n_channels)])
        assert rel_err(outs, ref_outs) < TOL, t
    for x, y in zip(weights(a), weights(r)):
        assert rel_err(x, y) < TOL
    assert a.contents.generation == r.contents.generation
    lib.rnn_batch_delete(batch)


def test_config4_multi_head_forward(gpu_lib, ref):
    """The shape of test/multi-text-6c34c563i73-h99-o3650.net (i73 / h99 /
    o3650 = 50 classes x 73 symbols, ReSQRT): forward over independent
    texts and the per-class cross entropy of
    rnn_char_multi_cross_entropy (charmodel-multi-predict.c:350-372)."""
    lib = gpu_lib
    n_texts, steps, n_classes, alpha = 6, 12, 50, 73
    shape = dict(input_size=alpha, hidden=99, output=n_classes * alpha, depth=5, seed=7,
                 activation=abi.RNN_RESQRT)
    r = make_net(ref, **shape)
    a = make_net(lib, **shape)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    rc = [ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n_texts)]
    ac = [lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n_texts)]
    arr_t = (abi.RecurNN_p * n_texts)(*ac)
    batch = lib.rnn_batch_new(arr_t, n_texts)
    assert batch
    rs = np.random.RandomState(7)
    text = rs.randint(0, alpha, size=(n_texts, steps + 1)).astype(np.uint8)
    ent_ref = np.zeros((n_texts, n_classes))
    ent_got = np.zeros((n_texts, n_classes))
    for t in range(steps):
        hot = np.ascontiguousarray(text[:, t])
        lib.rnn_batch_set_one_hot(batch, hot.ctypes.data_as(abi.u8_p))
        lib.rnn_batch_opinion(batch, 0.0)
        outs = np.zeros((n_texts, n_classes * alpha), dtype=np.float32)
        lib.rnn_batch_get_outputs(batch, fptr(outs))
        for j in range(n_texts):
            c = rc[j].contents
            x = arr(c.real_inputs, alpha)
            x[:] = 0
            x[text[j, t]] = 1.0
            o = arr(ref.rnn_opinion(rc[j], None, 0.0), n_classes * alpha).copy()
            assert rel_err(outs[j], o) < TOL
            for k in range(n_classes):
                for src, dst in ((o, ent_ref), (outs[j], ent_got)):
                    g = src[k * alpha:(k + 1) * alpha].astype(np.float64)
                    p = np.exp(g - g.max())
                    p /= p.sum()
                    dst[j, k] -= np.log2(max(p[text[j, t + 1]], 1e-30))
    assert rel_err(ent_got, ent_ref) < TOL
    lib.rnn_batch_delete(batch)


def test_config4_fixture_net_entropies_match_reference(gpu_lib, ref):
    """BASELINE.json configs[3] on the reference's own saved net
    (test/multi-text-6c34c563i73-h99-o3650.net, the CDB fixture), loaded by
    each side's rnn_load_net: the double[50] class entropies of
    rnn_char_multi_cross_entropy (charmodel-multi-predict.c:350-372) over
    synthetic texts, three ways - the reference; the reference's unmodified
    function linked against librecur_b200.so (its per-net rnn_opinion); and
    the array-of-nets calls with the reference's softmax on the outputs."""
    import oracle
    lib = gpu_lib
    if not (os.path.exists(oracle.FIXTURE_NET) and os.path.exists(oracle.CHARMULTI_B200)):
        pytest.skip("oracle/_ref fixtures not built")
    ours = oracle.load_charmulti_b200()
    n_texts, length, skip, alpha = 4, 300, 10, 73
    r = ref.rnn_load_net(oracle.FIXTURE_NET.encode())
    a = lib.rnn_load_net(oracle.FIXTURE_NET.encode())
    assert r and a
    n_classes = r.contents.output_size // alpha
    assert (r.contents.input_size, r.contents.hidden_size, n_classes) == (73, 99, 50)
    assert r.contents.activation == abi.RNN_RESQRT
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    rs = np.random.RandomState(7)
    text = rs.randint(0, alpha, size=(n_texts, length)).astype(np.uint8)
    want = np.zeros((n_texts, n_classes))
    per_net = np.zeros((n_texts, n_classes))
    dp = C.POINTER(C.c_double)
    for j in range(n_texts):
        rc = ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None)
        ref.rnn_char_multi_cross_entropy(rc, u8ptr(text[j]), length, alpha,
                                         want[j].ctypes.data_as(dp), skip)
        ac = lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None)
        ours.rnn_char_multi_cross_entropy(ac, u8ptr(text[j]), length, alpha,
                                          per_net[j].ctypes.data_as(dp), skip)
    assert want.min() > 0 and np.isfinite(want).all()
    assert rel_err(per_net, want) < TOL
    # the batch calls: all texts at once
    clones = [lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n_texts)]
    arr_t = (abi.RecurNN_p * n_texts)(*clones)
    batch = lib.rnn_batch_new(arr_t, n_texts)
    got = np.zeros((n_texts, n_classes))
    outs = np.zeros((n_texts, n_classes * alpha), dtype=np.float32)
    for t in range(length - 1):
        hot = np.ascontiguousarray(text[:, t])
        lib.rnn_batch_set_one_hot(batch, hot.ctypes.data_as(abi.u8_p))
        lib.rnn_batch_opinion(batch, 0.0)
        if t < skip:
            continue
        lib.rnn_batch_get_outputs(batch, fptr(outs))
        for j in range(n_texts):
            ref.ref_multi_entropy_step(fptr(outs[j]), n_classes, alpha, int(text[j, t + 1]),
                                       got[j].ctypes.data_as(dp))
    got /= (length - skip - 1)
    assert rel_err(got, want) < TOL
    lib.rnn_batch_delete(batch)


def test_config5_rnnca_cells_forward(gpu_lib, ref):
    """gstrnnca.c:805-820: one forward-only clone per cell, all sharing the
    trainers' weights; I35 / H51 / O3 (gstrnnca.h:13-51)."""
    lib = gpu_lib
    n_cells, frames = 600, 3
    shape = dict(input_size=35, hidden=51, output=3, depth=10, seed=11, lr=3e-3)
    r = make_net(ref, **shape)
    a = make_net(lib, **shape)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    ac = [lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n_cells)]
    batch = lib.rnn_batch_new((abi.RecurNN_p * n_cells)(*ac), n_cells)
    probe = [0, 1, 17, 299, 599]
    rc = {j: ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None) for j in probe}
    rs = np.random.RandomState(5)
    for f in range(frames):
        inputs = rs.random_sample((n_cells, 35)).astype(np.float32)
        lib.rnn_batch_set_inputs(batch, fptr(inputs))
        lib.rnn_batch_opinion(batch, 0.0)
        outs = np.zeros((n_cells, 3), dtype=np.float32)
        lib.rnn_batch_get_outputs(batch, fptr(outs))
        for j in probe:
            o = arr(ref.rnn_opinion(rc[j], fptr(inputs[j]), 0.0), 3)
            assert rel_err(outs[j], o) < TOL, (f, j)
    lib.rnn_batch_delete(batch)


def test_presynaptic_noise_draws_match_reference(gpu_lib, ref):
    """the per-stream Jenkins generator runs on the device;
    same seed, same draws, same noisy hidden state (recur-nn.c:120)."""
    lib = gpu_lib
    shape = dict(input_size=7, hidden=21, output=7, depth=4, seed=4, noise=0.1)
    r = make_net(ref, **shape)
    a = make_net(lib, **shape)
    rn = ref.rnn_new_training_set(r, 3)
    an = lib.rnn_new_training_set(a, 3)
    batch = lib.rnn_batch_new(an, 3)
    for t in range(4):
        hot = np.array([(t + j) % 7 for j in range(3)], dtype=np.uint8)
        lib.rnn_batch_advance(batch)
        lib.rnn_batch_set_one_hot(batch, hot.ctypes.data_as(abi.u8_p))
        lib.rnn_batch_opinion(batch, 0.1)
        for j in range(3):
            c = rn[j].contents
            ref.rnn_bptt_advance(rn[j])
            x = arr(c.real_inputs, 7)
            x[:] = 0
            x[hot[j]] = 1.0
            ref.rnn_opinion(rn[j], None, 0.1)
    lib.rnn_batch_pull(batch)
    for j in range(3):
        ca, cr = an[j].contents, rn[j].contents
        assert (ca.rng.a, ca.rng.b, ca.rng.c, ca.rng.d) == (cr.rng.a, cr.rng.b, cr.rng.c, cr.rng.d)
        assert rel_err(arr(ca.hidden_layer, ca.h_size), arr(cr.hidden_layer, cr.h_size)) < TOL
    # and through the per-net call
    b = make_net(lib, **shape)
    r2 = make_net(ref, **shape)
    for L, net in ((lib, b), (ref, r2)):
        c = net.contents
        for t in range(3):
            L.rnn_bptt_advance(net)
            x = arr(c.real_inputs, 7)
            x[:] = 0
            x[t] = 1.0
            L.rnn_opinion(net, None, 0.2)
    assert (b.contents.rng.a, b.contents.rng.d) == (r2.contents.rng.a, r2.contents.rng.d)
    assert rel_err(arr(b.contents.hidden_layer, 24), arr(r2.contents.hidden_layer, 24)) < TOL
    lib.rnn_batch_delete(batch)


def test_config4_multi_head_training_with_error_ranges(gpu_lib, ref):
    """text_train of charmodel-multi-predict.c:234-281 for one text: per
    step a softmax error on the target class's output range only
    (multi_softmax_error, 18-58), rnn_bptt_calc_deltas with RecurErrorRange,
    adagrad every batch_size steps.  Exercises the sparse top layer with its
    stale-row behaviour (recur-nn.c:156-196, 275-301) and ReSQRT backward."""
    lib = gpu_lib
    alpha, n_classes, target, batch_size, steps = 13, 5, 2, 4, 14
    shape = dict(input_size=alpha, hidden=31, output=n_classes * alpha, depth=6, seed=9,
                 lr=0.05, activation=abi.RNN_RESQRT)
    rs = np.random.RandomState(1)
    text = rs.randint(0, alpha, size=steps + 1)
    off = target * alpha
    ranges = (abi.RecurErrorRange * 2)()
    ranges[0].start = off & ~3
    ranges[0].len = ((off + alpha + 3) & ~3) - (off & ~3)
    ranges[1].start = -1
    res = []
    for L in (lib, ref):
        net = make_net(L, **shape)
        ih, ho = weights(net)
        ih *= 2.0   # livelier hidden layer: some units silent, some not
        L.rnn_set_momentum_values(net, 0.5)   # adagrad ballast
        c = net.contents
        b = c.bptt.contents
        countdown = batch_size
        trace = []
        for i in range(steps):
            L.rnn_bptt_advance(net)
            x = arr(c.real_inputs, alpha)
            x[:] = 0
            x[text[i]] = 1.0
            answer = arr(L.rnn_opinion(net, None, 0.0), c.o_size)
            err = arr(b.o_error, c.o_size)
            err[:] = 0
            seg = np.zeros(alpha, dtype=np.float32)
            ref.ref_softmax_best_guess(fptr(seg), fptr(np.ascontiguousarray(answer[off:off + alpha])), alpha)
            seg[text[i + 1]] += 1.0
            err[off:off + alpha] = seg
            countdown -= 1
            if countdown == 0:
                L.rnn_apply_learning(net, abi.RNN_ADAGRAD, b.momentum)
                countdown = batch_size
                L.rnn_bptt_calc_deltas(net, 0, ranges)
            else:
                L.rnn_bptt_calc_deltas(net, 1, ranges)
            trace.append((arr(b.ih_delta, c.ih_size).copy(), arr(b.ho_delta, c.ho_size).copy(),
                          b.ih_scale))
        res.append((trace, [w.copy() for w in weights(net)]))
    for i, (a, r) in enumerate(zip(res[0][0], res[1][0])):
        assert rel_err(a[0], r[0]) < TOL, ("ih_delta", i)
        assert rel_err(a[1], r[1]) < TOL, ("ho_delta", i)
        assert abs(a[2] - r[2]) < TOL
    for x, y in zip(res[0][1], res[1][1]):
        assert rel_err(x, y) < TOL


@pytest.mark.parametrize("n_nets", [6, 64])
def test_f1_char_classify_epoch_with_unlabelled_characters(gpu_lib, ref, n_nets):
    """rnn_char_classify_epoch (charmodel-classify.c:73-154): every stream runs
    forward on its character; only characters that carry a class are trained
    on, rnn_bptt_calc_deltas(n, j ? 1 : 0, NULL) — so when stream 0 sits a
    step out the deltas are NOT cleared that step (the reference's quirk,
    SURVEY.md §8a notes), which the masked batch call reproduces through its
    accumulate argument."""
    lib = gpu_lib
    alpha, classes, steps = 11, 3, 7
    NO_CLASS = 255
    shape = dict(input_size=alpha, hidden=67, output=classes, depth=5, seed=2, lr=0.01)
    rs = np.random.RandomState(n_nets)
    sym = rs.randint(0, alpha, size=(steps, n_nets)).astype(np.uint8)
    cls = rs.randint(0, classes, size=(steps, n_nets)).astype(np.uint8)
    cls[rs.random_sample((steps, n_nets)) < 0.3] = NO_CLASS
    cls[2, 0] = NO_CLASS      # stream 0 unlabelled: the stale-delta quirk
    cls[3, :] = NO_CLASS      # a step nobody trains on
    r = make_net(ref, **shape)
    a = make_net(lib, **shape)
    rn = ref.rnn_new_training_set(r, n_nets)
    an = lib.rnn_new_training_set(a, n_nets)
    batch = lib.rnn_batch_new(an, n_nets)
    for t in range(steps):
        for j in range(n_nets):
            c = rn[j].contents
            ref.rnn_bptt_advance(rn[j])
            x = arr(c.real_inputs, alpha)
            x[:] = 0
            x[sym[t, j]] = 1.0
            answer = ref.rnn_opinion(rn[j], None, 0.0)
            if cls[t, j] != NO_CLASS:
                ref.ref_softmax_best_guess(c.bptt.contents.o_error, answer, classes)
                arr(c.bptt.contents.o_error, c.o_size)[cls[t, j]] += 1.0
                ref.rnn_bptt_calc_deltas(rn[j], 1 if j else 0, None)
        ref.rnn_apply_learning(r, 0, 0.9)
        lib.rnn_batch_advance(batch)
        lib.rnn_batch_set_one_hot(batch, np.ascontiguousarray(sym[t]).ctypes.data_as(abi.u8_p))
        lib.rnn_batch_opinion(batch, 0.0)
        tgt = np.where(cls[t] == NO_CLASS, 0, cls[t]).astype(np.uint8)
        lib.rnn_batch_softmax_error(batch, tgt.ctypes.data_as(abi.u8_p), None, None)
        active = (cls[t] != NO_CLASS).astype(np.uint8)
        lib.rnn_batch_calc_deltas_masked(batch, 0 if active[0] else 1,
                                         active.ctypes.data_as(abi.u8_p))
        lib.rnn_apply_learning(a, 0, 0.9)
        for x, y in zip(weights(a), weights(r)):
            assert rel_err(x, y) < TOL, t
    lib.rnn_batch_pull(batch)
    for j in range(n_nets):
        assert an[j].contents.generation == rn[j].contents.generation
        assert rel_err(an[j].contents.bptt.contents.min_error_factor,
                       rn[j].contents.bptt.contents.min_error_factor) < TOL
    lib.rnn_batch_delete(batch)


@pytest.mark.parametrize("n", [3, 64])
def test_training_with_presynaptic_noise_matches_reference(gpu_lib, ref, n):
    """charmodel trains with presynaptic noise 0.1 (py-recur-text.c:443): noisy
    pad units fire too and must be handled as the reference does
    (recur-nn.c:120, 215-226, 335-337)."""
    lib = gpu_lib
    from helpers import markov_text, u8ptr
    text = markov_text(900, 11, seed=12)
    shape = dict(input_size=11, hidden=66, output=11, depth=5, seed=6, lr=0.01, noise=0.1)
    r = make_net(ref, **shape)
    a = make_net(lib, **shape)
    rn = ref.rnn_new_training_set(r, n)
    an = lib.rnn_new_training_set(a, n)
    batch = lib.rnn_batch_new(an, n)
    steps = 5
    ref.ref_multi_tap_train(rn, n, u8ptr(text), len(text), 0, steps, 0, 0.9, 0.0, None, None, None)
    lib.rnn_batch_text_upload(batch, u8ptr(text), len(text))
    lib.rnn_batch_text_train(batch, 0, steps, 0, 0.9, 0.0, None)
    lib.rnn_batch_pull(batch)
    for x, y in zip(weights(a), weights(r)):
        assert rel_err(x, y) < TOL
    for j in range(n):
        ca, cr = an[j].contents, rn[j].contents
        assert (ca.rng.a, ca.rng.d) == (cr.rng.a, cr.rng.d)
    lib.rnn_batch_delete(batch)


@pytest.mark.parametrize("cells_kernel", [True, False])
@pytest.mark.parametrize("edges,len_pos", [(1, 2), (0, 3)])
def test_f4_rnnca_frame_on_device(gpu_lib, ref, port, edges, len_pos, cells_kernel, monkeypatch):
    """SURVEY.md §8 f4, gstrnnca.c:805-830: a whole frame of the cellular
    automaton through rnn_batch_rnnca_frame (gather, forward, fast_sigmoid,
    bytes, all on the device) against the CPU: the oracle's restatement of
    fill_net_inputs (gstrnnca.c cannot be compiled, see oracle_rnn.c), the
    live reference's rnn_opinion and fast_sigmoid, UNIT_TO_BYTE.  Both device
    paths: the warp-per-cell kernel for tiny nets and gather + batch forward +
    emit."""
    lib = gpu_lib
    if not cells_kernel:
        monkeypatch.setenv("RECUR_B200_NO_CELLS", "1")
    W, Hh = 24, 16
    n = W * Hh
    # 17 luma and 8 chroma neighbours like the default pattern (gstrnnca.h:49-51)
    off_y = np.array([(dx, dy) for dy in range(-2, 3) for dx in range(-2, 3)
                      if abs(dx) + abs(dy) <= 2 or (abs(dx), abs(dy)) == (2, 2)][:17],
                     dtype=np.int32)
    off_c = np.array([(dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)],
                     dtype=np.int32)
    len_y, len_c = len(off_y), len(off_c)
    n_in = len_y + 2 * len_c + len_pos
    shape = dict(input_size=n_in, hidden=51, output=3, depth=10, seed=11, lr=3e-3)
    r = make_net(ref, **shape)
    a = make_net(lib, **shape)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    ac = [lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n)]
    rc = [ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n)]
    batch = lib.rnn_batch_new((abi.RecurNN_p * n)(*ac), n)
    rs = np.random.RandomState(9)
    frame = rs.randint(0, 256, size=3 * n).astype(np.uint8)
    u8p = C.POINTER(C.c_uint8)
    ip = C.POINTER(C.c_int)
    n_off = 0
    for f in range(3):
        got = np.zeros(3 * n, dtype=np.uint8)
        lib.rnn_batch_rnnca_frame(batch, frame.ctypes.data_as(u8p), got.ctypes.data_as(u8p), W, Hh,
                                  off_y.ctypes.data_as(ip), len_y, off_c.ctypes.data_as(ip), len_c,
                                  len_pos, edges)
        want = np.zeros(3 * n, dtype=np.uint8)
        inputs = np.zeros(n_in, dtype=np.float32)
        for cell in range(n):
            port.oracle_rnnca_fill_inputs(frame.ctypes.data_as(u8p), W, Hh, cell % W, cell // W,
                                          off_y.ctypes.data_as(ip), len_y,
                                          off_c.ctypes.data_as(ip), len_c, len_pos, edges,
                                          fptr(inputs))
            out = arr(ref.rnn_opinion(rc[cell], fptr(inputs), 0.0), 3)
            for i in range(3):
                want[i * n + cell] = port.oracle_rnnca_unit_to_byte(ref.ref_fast_sigmoid(float(out[i])))
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        # bytes are truncations of sigmoid * 255.9: a last-bit difference in the
        # float can move a value across an integer boundary, never further
        assert diff.max() <= 1, (f, diff.max())
        n_off += int((diff != 0).sum())
        frame = want  # both sides continue from the reference's frame
    assert n_off <= 0.01 * 3 * 3 * n, n_off
    lib.rnn_batch_delete(batch)


def _rnnca_pattern():
    # 17 luma and 8 chroma neighbours like the default pattern (gstrnnca.h:49-51)
    off_y = np.array([(dx, dy) for dy in range(-2, 3) for dx in range(-2, 3)
                      if abs(dx) + abs(dy) <= 2 or (abs(dx), abs(dy)) == (2, 2)][:17],
                     dtype=np.int32)
    off_c = np.array([(dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)],
                     dtype=np.int32)
    return off_y, off_c


def _rnnca_cpu_frame(L, port, clones, frame, W, Hh, off_y, off_c, len_pos, edges, n_in):
    """fill_frame (gstrnnca.c:805-830) on the CPU with library L's rnn_opinion
    and fast_sigmoid, the oracle's fill_net_inputs and UNIT_TO_BYTE."""
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    n = W * Hh
    want = np.zeros(3 * n, dtype=np.uint8)
    inputs = np.zeros(n_in, dtype=np.float32)
    for cell in range(n):
        port.oracle_rnnca_fill_inputs(frame.ctypes.data_as(u8p), W, Hh, cell % W, cell // W,
                                      off_y.ctypes.data_as(ip), len(off_y),
                                      off_c.ctypes.data_as(ip), len(off_c), len_pos, edges,
                                      fptr(inputs))
        out = arr(L.rnn_opinion(clones[cell], fptr(inputs), 0.0), 3)
        for i in range(3):
            want[i * n + cell] = port.oracle_rnnca_unit_to_byte(L.ref_fast_sigmoid(float(out[i])))
    return want


@pytest.mark.parametrize("kernel", ["tensor", "fma"])
@pytest.mark.parametrize("edges,len_pos,gain", [(1, 2, 1.0), (0, 3, 1.0), (0, 2, 12.0)])
def test_config5_device_cells_frame(gpu_lib, ref, ref_fast, port, edges, len_pos, gain, kernel,
                                    monkeypatch):
    """BASELINE configs[4] at its own shape: RnnCells (hidden state of every
    cell on the device, no host clone per pixel) through rnn_cells_rnnca_frame
    against fill_frame replayed on the CPU with BOTH builds of the reference.

    The bytes are truncations of fast_sigmoid(y) * 255.9 (gstrnnca.c:642), so
    a last-place difference in y moves a byte across an integer boundary, by
    one, never further.  The reference does that to itself: its shipped build
    (-Ofast -ffast-math -DVECTOR, what `ref_fast` is) and its IEEE build differ
    in some bytes by one; the tolerance here is that same +-1, and no more
    differing bytes than three times what the two reference builds show
    between themselves (plus a handful for tiny frames).

    gain 12 makes the hidden sums large enough for maybe_scale_inputs
    (recur-nn.c:68-81) to engage in every cell."""
    lib = gpu_lib
    monkeypatch.setenv("RECUR_B200_CELLS_FMA", "1" if kernel == "fma" else "0")
    W, Hh = 24, 17      # 408 cells: three full tiles of 128 and a ragged one
    n = W * Hh
    off_y, off_c = _rnnca_pattern()
    len_y, len_c = len(off_y), len(off_c)
    n_in = len_y + 2 * len_c + len_pos
    shape = dict(input_size=n_in, hidden=51, output=3, depth=10, seed=11, lr=3e-3)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    a = make_net(lib, **shape)
    nets, clones = {}, {}
    for name, L in (("strict", ref), ("fast", ref_fast)):
        nets[name] = make_net(L, **shape)
        clones[name] = [L.rnn_clone(nets[name], fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n)]
    for net in [a] + list(nets.values()):
        ih, ho = weights(net)
        ih *= gain
    cells = lib.rnn_cells_new(a, W, Hh)
    assert cells
    rs = np.random.RandomState(9)
    frame = rs.randint(0, 256, size=3 * n).astype(np.uint8)
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    ours_off = builds_off = 0
    clipped = 0
    for f in range(4 if gain > 1 else 3):
        got = np.zeros(3 * n, dtype=np.uint8)
        lib.rnn_cells_rnnca_frame(cells, frame.ctypes.data_as(u8p), got.ctypes.data_as(u8p),
                                  off_y.ctypes.data_as(ip), len_y, off_c.ctypes.data_as(ip), len_c,
                                  len_pos, edges)
        want = {k: _rnnca_cpu_frame(L, port, clones[k], frame, W, Hh, off_y, off_c, len_pos, edges,
                                    n_in) for k, L in (("strict", ref), ("fast", ref_fast))}
        d_ours = np.abs(got.astype(np.int32) - want["strict"].astype(np.int32))
        d_builds = np.abs(want["fast"].astype(np.int32) - want["strict"].astype(np.int32))
        assert d_ours.max() <= 1, (f, d_ours.max())
        assert d_builds.max() <= 1, (f, d_builds.max())
        ours_off += int((d_ours != 0).sum())
        builds_off += int((d_builds != 0).sum())
        # hidden state of a few cells against the strict clones
        for cell in (0, 1, W + 3, n - 1):
            h = np.zeros(52, dtype=np.float32)
            lib.rnn_cells_get_hidden(cells, cell, fptr(h))
            c = clones["strict"][cell].contents
            assert rel_err(h, arr(c.hidden_layer, c.h_size)) < TOL, (f, cell)
            x = arr(c.input_layer, c.i_size)
            clipped += bool(abs(x[0] - 1.0) > 1e-6)
        frame = want["strict"]
    assert ours_off <= 3 * builds_off + 12, (ours_off, builds_off)
    if gain > 1:
        assert clipped >= 4, clipped
    # rnn_cells_forget = fresh clones; rnn_cells_rnnca_run = the same frames chained
    lib.rnn_cells_forget(cells)
    start = rs.randint(0, 256, size=3 * n).astype(np.uint8)
    step = start.copy()
    for f in range(3):
        nxt = np.zeros_like(step)
        lib.rnn_cells_rnnca_frame(cells, step.ctypes.data_as(u8p), nxt.ctypes.data_as(u8p),
                                  off_y.ctypes.data_as(ip), len_y, off_c.ctypes.data_as(ip), len_c,
                                  len_pos, edges)
        step = nxt
    lib.rnn_cells_forget(cells)
    ran = np.zeros_like(step)
    lib.rnn_cells_rnnca_run(cells, start.ctypes.data_as(u8p), 3, ran.ctypes.data_as(u8p),
                            off_y.ctypes.data_as(ip), len_y, off_c.ctypes.data_as(ip), len_c,
                            len_pos, edges)
    assert np.array_equal(ran, step)
    lib.rnn_cells_delete(cells)


def test_config5_rnnca_trainer_step(gpu_lib, ref, port):
    """maybe_learn / train_net (gstrnnca.c:693-733) with the element's 200
    trainers (gstrnnca.h:36): clear the deltas; every trainer gathers its
    neighbourhood from the previous frame, runs forward (no rnn_bptt_advance:
    the element never calls it, the walk sees one live step and an empty
    ring), takes slope * (target - sigmoid) as the error and accumulates its
    deltas; then the weighted-momentum update with the soft start and
    rnn_condition_net.  On the device: the same calls as batch calls."""
    lib = gpu_lib
    W, Hh, n_tr, frames = 48, 32, 200, 12
    off_y, off_c = _rnnca_pattern()
    len_y, len_c, len_pos = len(off_y), len(off_c), 2
    n_in = len_y + 2 * len_c + len_pos
    flags = abi.RNN_NET_FLAG_STANDARD | abi.RNN_COND_USE_SCALE | abi.RNN_NET_FLAG_LOG_WEIGHT_SUM
    shape = dict(input_size=n_in, hidden=51, output=3, depth=10, seed=11, lr=3e-3, momentum=0.5,
                 flags=flags)
    rs = np.random.RandomState(21)
    pos = np.stack([rs.randint(2, W - 2, size=n_tr), rs.randint(2, Hh - 2, size=n_tr)], axis=1)
    yy, xx = np.mgrid[0:Hh, 0:W]
    movie = []
    for f in range(frames + 1):
        planes = [127 + 120 * np.sin(0.3 * xx + 0.2 * f + p) * np.cos(0.25 * yy - 0.1 * f * p)
                  for p in range(3)]
        movie.append(np.clip(np.stack(planes), 0, 255).astype(np.uint8).reshape(-1))
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    plane = W * Hh
    soft_start = 2000.0

    def gather(prev):
        x = np.zeros((n_tr, n_in), dtype=np.float32)
        for t in range(n_tr):
            port.oracle_rnnca_fill_inputs(prev.ctypes.data_as(u8p), W, Hh, int(pos[t, 0]),
                                          int(pos[t, 1]), off_y.ctypes.data_as(ip), len_y,
                                          off_c.ctypes.data_as(ip), len_c, len_pos, 1, fptr(x[t]))
        return x

    def errors(answers, now):
        err = np.zeros((n_tr, 4), dtype=np.float32)
        for t in range(n_tr):
            o = int(pos[t, 1]) * W + int(pos[t, 0])
            for i in range(3):
                a_ = np.float32(ref.ref_fast_sigmoid(float(answers[t, i])))
                target = np.float32(now[o + plane * i]) * np.float32(1.0 / 255.0)
                err[t, i] = a_ * (np.float32(1.0) - a_) * (target - a_)
        return err

    r = make_net(ref, **shape)
    a = make_net(lib, **shape)
    rn = ref.rnn_new_training_set(r, n_tr)
    an = lib.rnn_new_training_set(a, n_tr)
    batch = lib.rnn_batch_new(an, n_tr)
    for f in range(frames):
        prev, now = movie[f], movie[f + 1]
        x = gather(prev)
        # the reference, trainer by trainer
        ref.rnn_bptt_clear_deltas(r)
        answers = np.zeros((n_tr, 3), dtype=np.float32)
        for t in range(n_tr):
            c = rn[t].contents
            arr(c.real_inputs, n_in)[:] = x[t]
            answers[t] = arr(ref.rnn_opinion(rn[t], None, c.presynaptic_noise), 3)
        err_r = errors(answers, now)
        for t in range(n_tr):
            c = rn[t].contents
            arr(c.bptt.contents.o_error, c.o_size)[:] = err_r[t, :c.o_size]
            ref.rnn_bptt_calc_deltas(rn[t], 1, None)
        m = ref.rnn_calculate_momentum_soft_start(float(r.contents.generation),
                                                  r.contents.bptt.contents.momentum, soft_start)
        ref.rnn_apply_learning(r, abi.RNN_MOMENTUM_WEIGHTED, m)
        ref.rnn_condition_net(r)
        # the device, as batch calls
        lib.rnn_bptt_clear_deltas(a)
        lib.rnn_batch_set_inputs(batch, fptr(x))
        lib.rnn_batch_opinion(batch, 0.0)
        outs = np.zeros((n_tr, 3), dtype=np.float32)
        lib.rnn_batch_get_outputs(batch, fptr(outs))
        assert rel_err(outs, answers) < TOL, f
        err_a = np.ascontiguousarray(errors(outs, now)[:, :3])   # n x output_size
        lib.rnn_batch_set_errors(batch, fptr(err_a))
        lib.rnn_batch_calc_deltas(batch, 1)
        m = lib.rnn_calculate_momentum_soft_start(float(a.contents.generation),
                                                  a.contents.bptt.contents.momentum, soft_start)
        lib.rnn_apply_learning(a, abi.RNN_MOMENTUM_WEIGHTED, m)
        lib.rnn_condition_net(a)
        for p, q in zip(weights(a), weights(r)):
            assert rel_err(p, q) < TOL, f
        assert a.contents.generation == r.contents.generation
    lib.rnn_batch_delete(batch)


@pytest.mark.parametrize("activation,hidden,n_pos", [(abi.RNN_RESQRT, 51, 2), (abi.RNN_RECLIP20, 63, 0),
                                                     (abi.RNN_RELU, 20, 3)])
def test_device_cells_other_shapes(gpu_lib, ref, port, activation, hidden, n_pos):
    """RnnCells beyond rnnca's own shape: the other two activations
    (recur-nn.c:123-140), the widest net the kernel takes (63 hidden units: every
    column of the MMA tile in use), a narrow one (K steps skipped), no position
    terms / all three.  Tensor-core kernel against the reference cell by cell:
    hidden state of EVERY cell and the bytes."""
    lib = gpu_lib
    W, Hh = 20, 13      # 260 cells: two tiles and a ragged one
    n = W * Hh
    off_y, off_c = _rnnca_pattern()
    off_y = off_y[:9]
    len_y, len_c = len(off_y), len(off_c)
    n_in = len_y + 2 * len_c + n_pos
    shape = dict(input_size=n_in, hidden=hidden, output=3, depth=4, seed=3, lr=3e-3,
                 activation=activation)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    a, r = make_net(lib, **shape), make_net(ref, **shape)
    for net in (a, r):
        ih, ho = weights(net)
        ih *= 8.0 if activation == abi.RNN_RECLIP20 else 3.0   # the ceiling / the curve in play
    clones = [ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n)]
    cells = lib.rnn_cells_new(a, W, Hh)
    assert cells
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    frame = np.random.RandomState(4).randint(0, 256, size=3 * n).astype(np.uint8)
    h_size = a.contents.h_size
    for f in range(4):
        got = np.zeros(3 * n, dtype=np.uint8)
        lib.rnn_cells_rnnca_frame(cells, frame.ctypes.data_as(u8p), got.ctypes.data_as(u8p),
                                  off_y.ctypes.data_as(ip), len_y, off_c.ctypes.data_as(ip), len_c,
                                  n_pos, f % 2)
        want = _rnnca_cpu_frame(ref, port, clones, frame, W, Hh, off_y, off_c, n_pos, f % 2, n_in)
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and (d != 0).sum() <= 0.01 * d.size + 3, (f, d.max(), (d != 0).sum())
        worst = 0.0
        for cell in range(n):
            h = np.zeros(h_size, dtype=np.float32)
            lib.rnn_cells_get_hidden(cells, cell, fptr(h))
            c = clones[cell].contents
            worst = max(worst, rel_err(h, arr(c.hidden_layer, c.h_size)))
        assert worst < TOL, (f, worst)
        if activation == abi.RNN_RECLIP20 and f == 3:
            hs = np.concatenate([arr(cl.contents.hidden_layer, h_size) for cl in clones[:40]])
            assert (hs == 20.0).any()      # the ceiling was reached
        frame = want
    lib.rnn_cells_delete(cells)
    # shapes the kernel does not take are refused with a message, not mangled
    big = make_net(lib, input_size=n_in, hidden=100, output=3, depth=4, seed=3)
    assert not lib.rnn_cells_new(big, W, Hh)


def test_device_cells_rows_beyond_fp16(gpu_lib, ref, port):
    """Hidden values past FP16's range: the kernel stores such a cell's row
    divided by a power of two (CT_X_TOP, rb_cells.cu) and multiplies its sums
    back.  Weights x 3000 drive the hidden state of every cell past 65504 (the
    input soft clip, recur-nn.c:68-81, keeps it finite) - the reference, all
    in FP32, is matched to the usual tolerance."""
    lib = gpu_lib
    W, Hh = 16, 9
    n = W * Hh
    off_y, off_c = _rnnca_pattern()
    len_y, len_c, n_pos = len(off_y), len(off_c), 2
    n_in = len_y + 2 * len_c + n_pos
    shape = dict(input_size=n_in, hidden=51, output=3, depth=4, seed=5, lr=3e-3)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    a, r = make_net(lib, **shape), make_net(ref, **shape)
    for net in (a, r):
        ih, ho = weights(net)
        ih *= 3000.0
        ho *= 1e-5      # keep the outputs in the sigmoid's interesting range
    clones = [ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n)]
    cells = lib.rnn_cells_new(a, W, Hh)
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    frame = np.random.RandomState(8).randint(0, 256, size=3 * n).astype(np.uint8)
    biggest = 0.0
    for f in range(4):
        got = np.zeros(3 * n, dtype=np.uint8)
        lib.rnn_cells_rnnca_frame(cells, frame.ctypes.data_as(u8p), got.ctypes.data_as(u8p),
                                  off_y.ctypes.data_as(ip), len_y, off_c.ctypes.data_as(ip), len_c,
                                  n_pos, 1)
        want = _rnnca_cpu_frame(ref, port, clones, frame, W, Hh, off_y, off_c, n_pos, 1, n_in)
        worst = 0.0
        for cell in range(n):
            h = np.zeros(52, dtype=np.float32)
            lib.rnn_cells_get_hidden(cells, cell, fptr(h))
            want_h = arr(clones[cell].contents.hidden_layer, 52)
            worst = max(worst, rel_err(h, want_h))
            biggest = max(biggest, float(want_h.max()))
        assert worst < TOL, (f, worst)
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and (d != 0).sum() <= 0.02 * d.size + 3, (f, d.max(), (d != 0).sum())
        frame = want
    assert biggest > 65504.0, biggest
    lib.rnn_cells_delete(cells)
