"""The persistent BPTT walk of the tensor engine (rb_tc.cu, the headline kernel
of BASELINE.json configs[1]) against the UNMODIFIED reference compiled in
place (oracle/_ref, IEEE build), at the config's own net shape: hidden 1023,
depth 30, 64..160 synchronic streams.

A fresh net's ring is empty, so its first walks stop after a step or two
whatever the kernel does (rows of zeros mask every error).  The walks only
reach the depths of a long run once the ring is full.  Each case therefore
trains `warm` positions on the GPU, moves the complete training state into
the reference (helpers.transplant_training_set: teacher forcing, SURVEY.md
§8c), and then both sides take the SAME steps from the SAME state:

  plain     random weights as initialised:      walks stop unevenly at 13-14
  deep      weights x 1.5: the error does not die out, every walk runs all 30
  clipped   weights x 2.2: every stream's gradient is clipped (ih_scale != 1,
            recur-nn.c:393-402), walks end anywhere between 1 and 30
  (regimes found with the reference alone on the CPU; asserted below.)

Compared after each of the steps: weights, both deltas, every stream's
hidden layer, ih_scale, min_error_factor (1e-4 relative, max-norm) and the
executed depth of every stream (exact; the reference's comes from its log).
Each test asserts which kernel walked the ring.
"""
import ctypes as C

import numpy as np
import pytest

from recur_b200 import api
from helpers import (make_net, weights, arr, u8ptr, markov_text, rel_err,
                     transplant_training_set, reference_walk_depths)

pytestmark = pytest.mark.gpu
TOL = 1e-4
BIG = dict(input_size=42, hidden=1023, output=42, depth=30)


def run_case(lib, ref, tmp_path, shape, n, warm, steps, lr, boost, expect_kernel, engine=2):
    text = markov_text(20000, min(shape["input_size"], shape["output"]), seed=2)
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
    pos = lib.rnn_batch_text_train(batch, 0, warm, 0, 0.95, 2000.0, None)
    assert pos == warm
    lib.rnn_batch_pull(batch)
    transplant_training_set(gn, ref, rn, n)
    depth = shape["depth"]
    H, I = g.contents.h_size, g.contents.ih_size
    seen = dict(depths=[], clipped=0, x_sums=[])
    hs = g.contents.hidden_size
    for s in range(steps):
        want_depths = reference_walk_depths(
            ref, rn, n, depth,
            lambda: ref.ref_multi_tap_train(rn, n, u8ptr(text), len(text), warm + s, 1, 0,
                                            0.95, 2000.0, None, None, None),
            tmp_path)
        lib.rnn_batch_text_train(batch, warm + s, 1, 0, 0.95, 2000.0, None)
        assert lib.rnn_b200_last_walk_kernel().decode() == expect_kernel
        got_depths = (C.c_int32 * n)()
        lib.rnn_batch_bptt_depths(batch, got_depths)
        lib.rnn_batch_pull(batch)
        assert list(got_depths) == want_depths, s
        gb, rb = g.contents.bptt.contents, r.contents.bptt.contents
        assert rel_err(arr(gb.ih_delta, I), arr(rb.ih_delta, I)) < TOL, s
        assert rel_err(arr(gb.ho_delta, g.contents.ho_size),
                       arr(rb.ho_delta, g.contents.ho_size)) < TOL, s
        for x, y in zip(weights(g), weights(r)):
            assert rel_err(x, y) < TOL, s
        hg = np.stack([arr(gn[j].contents.hidden_layer, H) for j in range(n)])
        hr = np.stack([arr(rn[j].contents.hidden_layer, H) for j in range(n)])
        assert rel_err(hg, hr) < TOL, s
        if s + 1 < steps:   # what the next forward pass will find in its input row
            seen["x_sums"].append(2.0 + hr[:, 1:hs + 1].sum(axis=1))
        for field in ("ih_scale", "min_error_factor"):
            a = np.array([getattr(gn[j].contents.bptt.contents, field) for j in range(n)])
            b = np.array([getattr(rn[j].contents.bptt.contents, field) for j in range(n)])
            np.testing.assert_allclose(a, b, rtol=TOL, atol=0, err_msg="%s step %d" % (field, s))
            if field == "ih_scale":
                seen["clipped"] = max(seen["clipped"], int((b != 1).sum()))
        for j in range(n):
            assert gn[j].contents.bptt.contents.index == rn[j].contents.bptt.contents.index
        seen["depths"].append(want_depths)
    lib.rnn_batch_delete(batch)
    lib.rnn_delete_training_set(gn, n, 0)
    ref.rnn_delete_training_set(rn, n, 0)
    lib.rnn_b200_set_engine(0)
    return seen


@pytest.mark.parametrize("n", [64, 96, 160])
def test_persistent_chain_plain_regime_matches_reference(gpu_lib, ref, tmp_path, n):
    seen = run_case(gpu_lib, ref, tmp_path, BIG, n, warm=32, steps=2, lr=1e-6, boost=1.0,
                    expect_kernel="k_tc_chain_persistent")
    d = np.array(seen["depths"])
    assert d.min() >= 10 and d.max() < 30       # stopped by the error threshold
    assert len(np.unique(d)) > 1                # and not all at the same step


def test_persistent_chain_full_depth_matches_reference(gpu_lib, ref, tmp_path):
    seen = run_case(gpu_lib, ref, tmp_path, BIG, 64, warm=33, steps=2, lr=1e-6, boost=1.5,
                    expect_kernel="k_tc_chain_persistent")
    assert np.array(seen["depths"]).min() == 30


def test_persistent_chain_clipped_uneven_exits_match_reference(gpu_lib, ref, tmp_path):
    seen = run_case(gpu_lib, ref, tmp_path, BIG, 64, warm=31, steps=3, lr=1e-6, boost=2.2,
                    expect_kernel="k_tc_chain_persistent")
    d = np.array(seen["depths"])
    assert seen["clipped"] >= 10
    assert d.max() - d.min() >= 5


@pytest.mark.parametrize("shape,n,boost,lr", [
    (dict(input_size=42, hidden=199, output=42, depth=30), 96, 1.0, 2e-4),
    (dict(input_size=12, hidden=75, output=12, depth=8), 64, 3.0, 0.01),
])
def test_persistent_chain_on_small_nets_when_resident_walk_is_off(gpu_lib, ref, tmp_path,
                                                                  monkeypatch, shape, n,
                                                                  boost, lr):
    """Nets whose Wih fits one SM normally take k_walk_resident; with that
    switched off they run the persistent kernel's small-shape paths (one warp
    per row, fewer K splits)."""
    monkeypatch.setenv("RECUR_B200_NO_RESIDENT", "1")
    seen = run_case(gpu_lib, ref, tmp_path, shape, n, warm=shape["depth"] + 2, steps=2, lr=lr,
                    boost=boost, expect_kernel="k_tc_chain_persistent")
    if boost != 1.0:
        assert seen["clipped"] >= 10


def test_resident_walk_matches_reference_after_warmup(gpu_lib, ref, tmp_path):
    seen = run_case(gpu_lib, ref, tmp_path,
                    dict(input_size=42, hidden=199, output=42, depth=30), 96, warm=32, steps=2,
                    lr=2e-4, boost=1.0, expect_kernel="k_walk_resident")
    assert np.array(seen["depths"]).max() > 8


def test_forget_history_mid_run_on_the_tensor_engine(gpu_lib, ref):
    """gstclassify.c:1712-1714 forgets every channel's history (ring included)
    between segments.  The weight gradient reads the ring through operand
    planes; they must follow."""
    lib = gpu_lib
    n, shape = 64, dict(input_size=42, hidden=199, output=42, depth=12)
    text = markov_text(6000, 42, seed=6)
    g = make_net(lib, seed=2, lr=2e-4, **shape)
    r = make_net(ref, seed=2, lr=2e-4, **shape)
    gn = lib.rnn_new_training_set(g, n)
    rn = ref.rnn_new_training_set(r, n)
    lib.rnn_b200_set_engine(2)
    batch = lib.rnn_batch_new(gn, n)
    lib.rnn_batch_text_upload(batch, u8ptr(text), len(text))
    lib.rnn_batch_text_train(batch, 0, 14, 0, 0.95, 2000.0, None)
    ref.ref_multi_tap_train(rn, n, u8ptr(text), len(text), 0, 14, 0, 0.95, 2000.0, None, None, None)
    for j in range(n):
        lib.rnn_forget_history(gn[j], 1)
        ref.rnn_forget_history(rn[j], 1)
    lib.rnn_batch_text_train(batch, 14, 3, 0, 0.95, 2000.0, None)
    ref.ref_multi_tap_train(rn, n, u8ptr(text), len(text), 14, 3, 0, 0.95, 2000.0, None, None, None)
    lib.rnn_batch_pull(batch)
    I = g.contents.ih_size
    assert rel_err(arr(g.contents.bptt.contents.ih_delta, I),
                   arr(r.contents.bptt.contents.ih_delta, I)) < TOL
    for x, y in zip(weights(g), weights(r)):
        assert rel_err(x, y) < TOL
    lib.rnn_batch_delete(batch)
    lib.rnn_delete_training_set(gn, n, 0)
    ref.rnn_delete_training_set(rn, n, 0)
    lib.rnn_b200_set_engine(0)


@pytest.mark.parametrize("n,engine,kernel", [(5, 0, "k_walk_single"), (64, 2, "k_walk_resident")])
def test_input_soft_clip_in_batch_steps_matches_reference(gpu_lib, ref, tmp_path, n, engine,
                                                          kernel):
    """maybe_scale_inputs (recur-nn.c:68-81): weights x 10 blow the hidden sum
    past 16 * i_size, the input row is scaled down in the ring (and BPTT walks
    the scaled rows).  FMA engine at 5 streams, tensor engine at 64."""
    shape = dict(input_size=12, hidden=75, output=12, depth=6)
    seen = run_case(gpu_lib, ref, tmp_path, shape, n, warm=7, steps=3, lr=1e-5, boost=10.0,
                    expect_kernel=kernel, engine=engine)
    i_size = (75 + 12 + 1 + 3) // 4 * 4
    assert max(x.max() for x in seen["x_sums"]) > 16 * i_size     # the clip really ran
