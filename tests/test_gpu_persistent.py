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

Every compared step starts from transplanted, bit-identical state (the GPU
keeps its own trajectory; the reference is re-seeded from it each time).
Compared per step: the reference's own log keys of every stream (recur-nn.c:
415-421,766-770: depth, ih_scale, min_error_threshold, min_error_factor,
cum_error, scaled_error, top_error_*; five printed digits), weights, both
deltas, every stream's hidden layer, ih_scale and min_error_factor (1e-4
relative, max-norm).  Executed depth is an integer and must be equal, with one
allowance: the walk ends where error_sum crosses a threshold (recur-nn.c:387),
and fp32 sums taken in a different order differ by a few 1e-5 after a dozen
chained steps (the FMA engine, exact fp32, shows the same against the
reference), so a stream whose deciding sum lies within 1e-3 of its threshold
may end one step apart - at most 1 % of the streams of a step, each proven
from the logged numbers; steps where that happens compare the untouched
streams only, and most steps of a case must be free of it.  The same goes
for a hidden unit whose pre-activation is zero to within rounding: its value
agrees, the top layer's `hidden != 0` mask does not (checked from the
hidden layers themselves).
Each test asserts which kernel walked the ring.
"""
import ctypes as C

import numpy as np
import pytest

from recur_b200 import api
from helpers import (make_net, weights, arr, u8ptr, markov_text, rel_err,
                     transplant_training_set, reference_walk_logs, executed_depth)

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
    try:
        return _compare_steps(lib, ref, tmp_path, n, warm, steps, expect_kernel, text, g, r, gn,
                              rn, batch, depth, H, I)
    finally:
        lib.rnn_batch_delete(batch)
        lib.rnn_delete_training_set(gn, n, 0)
        ref.rnn_delete_training_set(rn, n, 0)
        lib.rnn_b200_set_engine(0)


def _compare_steps(lib, ref, tmp_path, n, warm, steps, expect_kernel, text, g, r, gn, rn, batch,
                   depth, H, I):
    seen = dict(depths=[], clipped=0, x_sums=[], flip_steps=0)
    hs = g.contents.hidden_size
    for s in range(steps):
        logs = reference_walk_logs(
            ref, rn, n,
            lambda: ref.ref_multi_tap_train(rn, n, u8ptr(text), len(text), warm + s, 1, 0,
                                            0.95, 2000.0, None, None, None),
            tmp_path)
        want_depths = [executed_depth(l["depth"], depth) for l in logs]
        lib.rnn_batch_text_train(batch, warm + s, 1, 0, 0.95, 2000.0, None)
        assert lib.rnn_b200_last_walk_kernel().decode() == expect_kernel
        got = (api.RnnBatchBpttLog * n)()
        lib.rnn_batch_bptt_log(batch, got)
        lib.rnn_batch_pull(batch)
        hg = np.stack([arr(gn[j].contents.hidden_layer, H) for j in range(n)])
        hr = np.stack([arr(rn[j].contents.hidden_layer, H) for j in range(n)])
        assert rel_err(hg, hr) < TOL, s
        # ReLU at zero: a unit whose sum is zero to within rounding is off on
        # one side and barely on on the other.  The values agree; the top
        # layer's `hidden != 0` mask (recur-nn.c:221) does not, and that
        # stream's error differs by the unit's share from there on.
        masked = [j for j in range(n) if ((hg[j] == 0) != (hr[j] == 0)).any()]
        for j in masked:
            d = (hg[j] == 0) != (hr[j] == 0)
            assert np.maximum(np.abs(hg[j][d]), np.abs(hr[j][d])).max() < 1e-5 * np.abs(hr[j]).max()
        # the reference's own log keys (five printed digits) for every stream
        for key, field in (("top_error_raw", "top_error_raw"),
                           ("top_error_scaled", "top_error_scaled"),
                           ("min_error_threshold", "min_error_threshold")):
            a = np.array([getattr(got[j], field) for j in range(n) if j not in masked])
            b = np.array([logs[j][key] for j in range(n) if j not in masked])
            # 1e-3: a confident stream's error 1 - p cancels (p near 1), so the
            # forward pass's 1e-6 comes back multiplied by p / (1 - p)
            np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-6 * np.abs(b).max(),
                                       err_msg="%s step %d" % (key, s))
        flips = [j for j in range(n) if got[j].n_steps != want_depths[j]]
        for j in flips:
            lo, hi = logs[j]["min_error_threshold"], 2.0 * logs[j]["top_error_scaled"] + 1.0
            # the side that stopped first holds the sum that decided
            if got[j].n_steps < want_depths[j]:
                es = got[j].error_sum
            else:
                es = logs[j]["scaled_error"] / logs[j]["ih_scale"]
            near = min(abs(es - lo) / lo, abs(es - hi) / hi)
            assert abs(got[j].n_steps - want_depths[j]) == 1 and near < 1e-3, \
                (s, j, got[j].n_steps, want_depths[j], es, lo, hi)
        flips = sorted(set(flips) | set(masked))
        assert len(flips) <= max(1, n // 50), (s, flips)
        seen["flip_steps"] += bool(flips)
        keep = [j for j in range(n) if j not in flips]
        for key, field in (("depth", "depth"), ("ih_scale", "ih_scale"),
                           ("min_error_factor", "min_error_factor"),
                           ("cum_error", "cum_error"), ("scaled_error", "scaled_error")):
            a = np.array([getattr(got[j], field) for j in keep])
            b = np.array([logs[j][key] for j in keep])
            # error_sum closes a chain of up to 30 dependent products: its own
            # tolerance (the FMA engine, exact fp32, is 2e-5..3e-4 from the reference)
            chained = key in ("scaled_error", "cum_error", "ih_scale")
            np.testing.assert_allclose(a, b, rtol=1e-3 if chained else 2e-4,
                                       atol=1e-5 * max(1.0, np.abs(b).max()) if chained else 0,
                                       err_msg="%s step %d" % (key, s))
        if s + 1 < steps:   # what the next forward pass will find in its input row
            seen["x_sums"].append(2.0 + hr[:, 1:hs + 1].sum(axis=1))
        gb, rb = g.contents.bptt.contents, r.contents.bptt.contents
        assert rel_err(arr(gb.ho_delta, g.contents.ho_size),
                       arr(rb.ho_delta, g.contents.ho_size)) < TOL, s
        if not flips:
            assert rel_err(arr(gb.ih_delta, I), arr(rb.ih_delta, I)) < TOL, s
            for x, y in zip(weights(g), weights(r)):
                assert rel_err(x, y) < TOL, s
        a = np.array([gn[j].contents.bptt.contents.min_error_factor for j in keep])
        b = np.array([rn[j].contents.bptt.contents.min_error_factor for j in keep])
        np.testing.assert_allclose(a, b, rtol=TOL, atol=0, err_msg="mef step %d" % s)
        a = np.array([gn[j].contents.bptt.contents.ih_scale for j in keep])
        b = np.array([rn[j].contents.bptt.contents.ih_scale for j in keep])
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-5, err_msg="ih_scale step %d" % s)
        seen["clipped"] = max(seen["clipped"], int((b != 1).sum()))
        for j in range(n):
            assert gn[j].contents.bptt.contents.index == rn[j].contents.bptt.contents.index
        if s + 1 < steps:
            transplant_training_set(gn, ref, rn, n)   # teacher forcing: re-seed the reference
        seen["depths"].append(want_depths)
    assert seen["flip_steps"] * 2 <= steps, seen["flip_steps"]
    return seen


@pytest.mark.parametrize("n", [64, 96, 160])
def test_persistent_chain_plain_regime_matches_reference(gpu_lib, ref, tmp_path, n):
    seen = run_case(gpu_lib, ref, tmp_path, BIG, n, warm=32, steps=4, lr=1e-6, boost=1.0,
                    expect_kernel="k_tc_chain_persistent")
    d = np.array(seen["depths"])
    assert d.min() >= 10 and d.max() < 30       # stopped by the error threshold
    assert len(np.unique(d)) > 1                # and not all at the same step


def test_persistent_chain_full_depth_matches_reference(gpu_lib, ref, tmp_path):
    seen = run_case(gpu_lib, ref, tmp_path, BIG, 64, warm=33, steps=3, lr=1e-6, boost=1.5,
                    expect_kernel="k_tc_chain_persistent")
    assert np.array(seen["depths"]).min() == 30


def test_persistent_chain_clipped_uneven_exits_match_reference(gpu_lib, ref, tmp_path):
    seen = run_case(gpu_lib, ref, tmp_path, BIG, 64, warm=31, steps=6, lr=1e-6, boost=2.2,
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
