#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the UNMODIFIED
reference compiled in place (oracle/_ref/librecur_ref_strict.so, the IEEE
-O2 -fno-fast-math build; see oracle/Makefile).  Run in the build container,
where /root/reference exists:

    python tests/golden/make_golden.py

Outputs (committed):
    recur_golden.npz     scalar helpers, RNG draws, initial weights and a
                         40-step synchronic training trace of a small net
    ref_saved_small.net  that net as written by the reference's rnn_save_net
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from recur_b200 import abi  # noqa: E402
from helpers import make_net, arr, weights, markov_text, u8ptr, fptr  # noqa: E402

SMALL = dict(input_size=7, hidden=13, output=7, depth=6, seed=3, lr=0.02)
N_STREAMS = 3
N_STEPS = 40


HOT = dict(SMALL, lr=0.1)   # with weights doubled: ih_scale clips 13 times
HOT_BOOST = 2.0


def trace_training(lib, n_streams=N_STREAMS, n_steps=N_STEPS, cfg=SMALL,
                   momentum=0.9, style=abi.RNN_MOMENTUM_WEIGHTED, text=None,
                   boost=1.0):
    """The multi-tap loop of charmodel-predict.c:293-311 driven from Python,
    recording what each stage leaves behind."""
    net = make_net(lib, **cfg)
    if boost != 1.0:
        ih, ho = weights(net)
        ih *= boost
        ho *= boost
    nets = lib.rnn_new_training_set(net, n_streams)
    n = net.contents
    if text is None:
        text = markov_text(400, cfg["input_size"], seed=5)
    length = len(text)
    spacing = (length - 1) // n_streams
    out = {k: [] for k in ("hidden", "output", "o_error", "ih_scale", "mef",
                           "ih_delta", "ho_delta", "ih_weights", "ho_weights",
                           "err", "winner")}
    ih0, ho0 = weights(net)
    out["ih_weights0"] = ih0.copy()
    out["ho_weights0"] = ho0.copy()
    for i in range(n_steps):
        rows = {k: [] for k in ("hidden", "output", "o_error", "ih_scale", "mef",
                                "err", "winner")}
        for j in range(n_streams):
            nj = nets[j]
            c = nj.contents
            off = i + j * spacing
            if off >= length - 1:
                off -= length - 1
            lib.rnn_bptt_advance(nj)
            inputs = arr(c.real_inputs, c.input_size)
            inputs[:] = 0
            inputs[text[off]] = 1.0
            answer = lib.rnn_opinion(nj, None, 0.0)
            err = arr(c.bptt.contents.o_error, c.o_size)
            winner = lib.ref_softmax_best_guess(c.bptt.contents.o_error, answer,
                                                c.output_size) \
                if hasattr(lib, "ref_softmax_best_guess") else None
            err[text[off + 1]] += 1.0
            rows["hidden"].append(arr(c.hidden_layer, c.h_size).copy())
            rows["output"].append(arr(c.output_layer, c.o_size).copy())
            rows["o_error"].append(err.copy())
            rows["err"].append(float(err[text[off + 1]]))
            rows["winner"].append(-1 if winner is None else int(winner))
            lib.rnn_bptt_calc_deltas(nj, 1 if j else 0, None)
            rows["ih_scale"].append(c.bptt.contents.ih_scale)
            rows["mef"].append(c.bptt.contents.min_error_factor)
        b = n.bptt.contents
        out["ih_delta"].append(arr(b.ih_delta, n.ih_size).copy())
        out["ho_delta"].append(arr(b.ho_delta, n.ho_size).copy())
        lib.rnn_apply_learning(net, style, momentum)
        ih, ho = weights(net)
        out["ih_weights"].append(ih.copy())
        out["ho_weights"].append(ho.copy())
        for k, v in rows.items():
            out[k].append(np.array(v))
    res = {k: np.array(v) for k, v in out.items()}
    res["text"] = text
    return res, net, nets


def main():
    if not os.path.exists("/root/reference/recur-nn.c"):
        sys.exit("needs /root/reference")
    oracle.build(ref=True, port=False)
    ref = oracle.load_ref(strict=True)
    g = {}
    # fast_expf over the range softmax can feed it (badmaths.h:14-29)
    xs = np.concatenate([np.linspace(-62, 52, 229), np.array([0.0, 0.2, -0.2, 0.19999999, 1e-9])]).astype(np.float32)
    g["expf_x"] = xs
    g["expf_y"] = np.array([ref.ref_fast_expf(float(x)) for x in xs], dtype=np.float32)
    # soft_clip (recur-nn-helpers.h:104-113)
    sc = np.array([[1.0, 2.0], [5.0, 2.0], [100.0, 3.0], [17.0, 16.0], [3.3, 0.0]], dtype=np.float32)
    g["softclip_in"] = sc
    g["softclip_out"] = np.array([ref.ref_soft_clip(float(a), float(b)) for a, b in sc], dtype=np.float32)
    # softmax_best_guess incl. both clamp branches (badmaths.h:71-141)
    rs = np.random.RandomState(11)
    cases = [rs.randn(42) * 3, rs.randn(42) * 3 + 70, rs.randn(42) * 3 - 90,
             np.linspace(-80, 60, 42), rs.randn(5), np.zeros(7)]
    for k, y in enumerate(cases):
        y = y.astype(np.float32)
        e = np.zeros_like(y)
        w = ref.ref_softmax_best_guess(fptr(e), fptr(y), len(y))
        g["softmax_y_%d" % k] = y
        g["softmax_e_%d" % k] = e
        g["softmax_w_%d" % k] = np.int32(w)
    g["softmax_n"] = np.int32(len(cases))
    # generator draws (recur-rng.h)
    draws = []
    for seed in (1, 2, 11, 12345678901234567):
        ctx = abi.RandCtx()
        ref.ref_init_rand64(C.byref(ctx), seed)
        d64 = [ref.ref_rand64(C.byref(ctx)) for _ in range(8)]
        dd = [ref.ref_rand_double(C.byref(ctx)) for _ in range(4)]
        dg = [ref.ref_cheap_gaussian_noise(C.byref(ctx)) for _ in range(4)]
        draws.append((seed, d64, dd, dg))
    g["rng_seeds"] = np.array([d[0] for d in draws], dtype=np.uint64)
    g["rng_u64"] = np.array([d[1] for d in draws], dtype=np.uint64)
    g["rng_double"] = np.array([d[2] for d in draws], dtype=np.float64)
    g["rng_gauss"] = np.array([d[3] for d in draws], dtype=np.float32)
    # a training trace
    tr, net, nets = trace_training(ref)
    for k, v in tr.items():
        g["trace_" + k] = v
    ref.rnn_save_net(net, os.path.join(HERE, "ref_saved_small.net").encode(), 0)
    tr2, _, _ = trace_training(ref, cfg=HOT, boost=HOT_BOOST)
    for k, v in tr2.items():
        g["hot_" + k] = v
    print("hot trace: ih_scale min", tr2["ih_scale"].min(), "clipped calls",
          int((tr2["ih_scale"] != 1).sum()))
    np.savez_compressed(os.path.join(HERE, "recur_golden.npz"), **g)
    print("wrote", os.path.join(HERE, "recur_golden.npz"),
          "ih_scale range", tr["ih_scale"].min(), tr["ih_scale"].max(),
          "mef", tr["mef"].min(), tr["mef"].max())


if __name__ == "__main__":
    main()
