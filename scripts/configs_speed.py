"""BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py): stream-steps/s of
this library on the GPU next to the compiled reference (oracle/_ref, its own
flags) on ONE host core of the same box, same loops as tests/test_gpu_configs.py
(the GStreamer elements themselves cannot be built).  The
reference loops run through ctypes: a few microseconds of Python per call are
included in their times (<= 10 %).  Prints a markdown table."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np
import oracle
from recur_b200 import api, abi
from helpers import make_net, u8ptr, fptr, arr, markov_text, STD_FLAGS

lib = None   # this library (needs a GPU) and the compiled reference, loaded on demand:
ref = None   # bench.py --impl reference runs the functions below with ours=False
rows = []


def load(ours=True):
    global lib, ref
    if ours and lib is None:
        lib = api.load_library()
    if ref is None:
        ref = oracle.load_ref(strict=False)


def softmax_err(out, target):
    e = np.exp(out - out.max(axis=-1, keepdims=True))
    p = e / e.sum(axis=-1, keepdims=True)
    err = -p
    np.put_along_axis(err, np.asarray(target)[..., None], np.take_along_axis(err, np.asarray(target)[..., None], -1) + 1.0, -1)
    return err.astype(np.float32)


# ---- config 1: default text-predict, one net ---------------------------------
def config1(ours=True, theirs=True):
    load(ours)
    name = "1 text-predict default (1 net, H199, depth 30, per-net API)"
    text = markov_text(20000, 42, seed=6)
    shape = dict(input_size=42, hidden=199, output=42, depth=30, seed=1, lr=1e-3)
    steps = 2000
    ref_rate = None
    if theirs:
        r = make_net(ref, **shape)
        t_ref = ref.ref_single_net_train(r, u8ptr(text), len(text), 0, steps, 0.95, 2000.0, 1, None, None, None)
        ref_rate = steps / t_ref
    if not ours:
        rows.append((name, "chars/s", None, ref_rate))
        return rows[-1]
    a = make_net(lib, **shape)
    c = a.contents
    o_err = c.bptt.contents.o_error

    def run(n0, n1):
        for i in range(n0, n1):
            c.bptt.contents.momentum = lib.rnn_calculate_momentum_soft_start(c.generation, 0.95, 2000.0)
            lib.rnn_bptt_advance(a)
            x = arr(c.real_inputs, c.input_size)
            x[:] = 0
            x[text[i]] = 1.0
            ans = lib.rnn_opinion(a, None, 0.0)
            ref.ref_softmax_best_guess(o_err, ans, c.output_size)
            o_err[text[i + 1]] += 1.0
            lib.rnn_bptt_calculate(a, 1)
    run(0, 100)
    t0 = time.perf_counter(); run(100, 1100); t = time.perf_counter() - t0
    rows.append((name, "chars/s", 1000 / t, ref_rate))
    return rows[-1]


# ---- config 3: classify, 256 channels -----------------------------------------
def config3(ours=True, theirs=True):
    load(ours)
    name = "3 classify training (256 channels, 32 features, H199, depth 30, Nesterov)"
    B, F, Hn, classes, depth = 256, 32, 199, 4, 30
    rs = np.random.RandomState(3)
    kw = dict(input_size=F, hidden=Hn, output=classes, depth=depth, seed=11, lr=1e-6, flags=STD_FLAGS)
    if theirs:
        r = make_net(ref, **kw)
        rn = ref.rnn_new_training_set(r, B)
    if ours:
        a = make_net(lib, **kw)
        an = lib.rnn_new_training_set(a, B)
        batch = lib.rnn_batch_new(an, B)

    def run_ours(chunks):
        outs = np.zeros((B, classes), dtype=np.float32)
        for t in range(chunks):
            feats = np.log1p(rs.random_sample((B, F)) * 400).astype(np.float32)
            targets = rs.randint(0, classes, size=B)
            lib.rnn_batch_set_inputs(batch, fptr(feats))
            lib.rnn_batch_opinion(batch, 0.0)
            lib.rnn_batch_get_outputs(batch, fptr(outs))
            errs = np.ascontiguousarray(softmax_err(outs, targets))
            lib.rnn_batch_set_errors(batch, fptr(errs))
            lib.rnn_batch_calc_deltas(batch, 0)
            lib.rnn_batch_advance(batch)
            lib.rnn_apply_learning(a, abi.RNN_MOMENTUM_NESTEROV, 0.9)
            lib.rnn_condition_net(a)

    def run_theirs(chunks):
        for t in range(chunks):
            feats = np.log1p(rs.random_sample((B, F)) * 400).astype(np.float32)
            targets = rs.randint(0, classes, size=B)
            ref.rnn_bptt_clear_deltas(r)
            for j in range(B):
                cj = rn[j].contents
                out = ref.rnn_opinion(rn[j], fptr(feats[j]), 0.0)
                err = arr(cj.bptt.contents.o_error, cj.o_size)
                err[:classes] = softmax_err(arr(out, classes), targets[j])
                ref.rnn_bptt_calc_deltas(rn[j], 1, None)
                ref.rnn_bptt_advance(rn[j])
            ref.rnn_apply_learning(r, abi.RNN_MOMENTUM_NESTEROV, 0.9)
            ref.rnn_condition_net(r)
    v = rv = None
    if ours:
        run_ours(40)
        t0 = time.perf_counter(); run_ours(200); t = time.perf_counter() - t0
        v = 200 * B / t
    if theirs:
        run_theirs(2)
        t0 = time.perf_counter(); run_theirs(6); tr = time.perf_counter() - t0
        rv = 6 * B / tr
    rows.append((name, "channel-frames/s", v, rv))
    return rows[-1]


# ---- config 4: multi-head charmodel forward ------------------------------------
def config4(ours=True, theirs=True):
    load(ours)
    name = "4 charmodel multi-head forward (64 texts, i73/h99/o3650, ReSQRT, outputs to host)"
    n_texts, n_classes, alpha = 64, 50, 73
    shape = dict(input_size=alpha, hidden=99, output=n_classes * alpha, depth=5, seed=7, activation=abi.RNN_RESQRT)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    rs = np.random.RandomState(7)
    text = rs.randint(0, alpha, size=(n_texts, 600)).astype(np.uint8)
    outs = np.zeros((n_texts, n_classes * alpha), dtype=np.float32)
    v = rv = None
    if ours:
        a = make_net(lib, **shape)
        ac = [lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n_texts)]
        batch = lib.rnn_batch_new((abi.RecurNN_p * n_texts)(*ac), n_texts)

        def run_ours(t0s, t1s):
            for t in range(t0s, t1s):
                hot = np.ascontiguousarray(text[:, t])
                lib.rnn_batch_set_one_hot(batch, hot.ctypes.data_as(abi.u8_p))
                lib.rnn_batch_opinion(batch, 0.0)
                lib.rnn_batch_get_outputs(batch, fptr(outs))
        run_ours(0, 50)
        t0 = time.perf_counter(); run_ours(50, 550); t = time.perf_counter() - t0
        v = 500 * n_texts / t
    if theirs:
        r = make_net(ref, **shape)
        rc = [ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(4)]
        t0 = time.perf_counter()
        for tt in range(500):
            for j in range(4):
                cj = rc[j].contents
                x = arr(cj.real_inputs, alpha)
                x[:] = 0
                x[text[j, tt]] = 1.0
                ref.rnn_opinion(rc[j], None, 0.0)
        tr = time.perf_counter() - t0
        rv = 500 * 4 / tr
    rows.append((name, "text-chars/s", v, rv))
    return rows[-1]


# ---- config 5: rnnca cells ------------------------------------------------------
def config5():
    load()
    n = 144 * 96
    a = make_net(lib, input_size=35, hidden=51, output=3, depth=10, seed=11, lr=3e-3)
    r = make_net(ref, input_size=35, hidden=51, output=3, depth=10, seed=11, lr=3e-3)
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    cells = [lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n)]
    batch = lib.rnn_batch_new((abi.RecurNN_p * n)(*cells), n)
    rs = np.random.RandomState(5)
    inputs = rs.random_sample((n, 35)).astype(np.float32)
    outs = np.zeros((n, 3), dtype=np.float32)

    def ours(frames):
        for f in range(frames):
            lib.rnn_batch_set_inputs(batch, fptr(inputs))
            lib.rnn_batch_opinion(batch, 0.0)
            lib.rnn_batch_get_outputs(batch, fptr(outs))
    ours(5)
    t0 = time.perf_counter(); ours(200); t = time.perf_counter() - t0
    rc = [ref.rnn_clone(r, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(200)]
    t0 = time.perf_counter()
    for rep in range(20):
        for j in range(200):
            ref.rnn_opinion(rc[j], fptr(inputs[j]), 0.0)
    tr = time.perf_counter() - t0
    rows.append(("5 rnnca forward (144 x 96 = 13,824 cells, i35/h51/o3, frames in and out of host memory)", "cell-steps/s", 200 * n / t, 20 * 200 / tr))


if __name__ == "__main__":
    for fn in (config1, config3, config4, config5):
        fn()
    print("| config | unit | this library (1 B200) | reference (1 host core) | ratio |")
    print("|---|---|---:|---:|---:|")
    for name, unit, ours_v, ref_v in rows:
        print("| %s | %s | %.3g | %.3g | %.1f |" % (name, unit, ours_v, ref_v, ours_v / ref_v))
