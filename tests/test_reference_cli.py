"""The drop-in boundary exercised by the caller it was made for: the
reference's OWN command-line trainer - unmodified text-predict.c,
charmodel-predict.c, charmodel-init.c and ccan/opt (reference Makefile:188-191)
- built by oracle/Makefile (target `cli`) twice: against the reference's RNN
core, and against this repo's librecur_b200.so with nothing between them but
the linker.  BASELINE.json configs[0]: text-predict on
test-images/erewhon.txt with default options (hidden 199, depth 30, single net:
the rnn_bptt_calculate path, recur-nn.c:999-1019), --rng-seed=1.

Compared: what the reference itself logs for every generation
(recur-nn.c:415-448,766-771: depth, ih_scale, min_error_threshold,
min_error_factor, cum_error, hidden_sum, hidden_magnitude, hidden_zeros,
top_error_raw/scaled, error_sum) and per report interval
(charmodel-predict.c:340-376: t_entropy, t_error, accuracy).  A single stream
trained free-running is chaotic - every 1e-7 of fp32 summation order is
amplified as the run goes on - so the step-by-step comparison covers the first
generations, the rest of the run is compared through its statistics and the
learning curve.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "oracle", "_ref", "cli")
REF_BIN = os.path.join(CLI, "text-predict-ref")
OUR_BIN = os.path.join(CLI, "text-predict-b200")
TEXT = os.path.join(CLI, "erewhon.txt")


def need_cli():
    if not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN) and os.path.exists(TEXT)):
        if os.path.exists("/root/reference/text-predict.c"):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "cli"], check=True)
        else:
            pytest.skip("oracle/_ref/cli not built and /root/reference absent")


def run_cli(binary, workdir, stop_after, extra=()):
    os.makedirs(os.path.join(workdir, "nets"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "images"), exist_ok=True)
    log = os.path.join(workdir, "run.log")
    cmd = [binary, "--text-file=" + TEXT, "--stop-after=%d" % stop_after, "--rng-seed=1",
           "--no-save-net", "--log-file=" + log, "-q"] + list(extra)
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, timeout=900)
    return r, log


def parse_log(path):
    """generation-indexed records of the per-step keys, and the list of
    report-interval records."""
    steps, reports, cur = [], [], {}
    report_keys = ("t_entropy", "t_error", "v_entropy", "accuracy", "per_second", "momentum",
                   "learn-rate")
    rep = {}
    for line in open(path):
        key, _, val = line.strip().partition(" ")
        if not val:
            continue
        if key == "generation":
            if cur:
                steps.append(cur)
            cur = {}
            if rep:
                reports.append(rep)
                rep = {}
        elif key in report_keys:
            rep[key] = float(val)
        else:
            cur[key] = float(val)
    if rep:
        reports.append(rep)
    return steps, reports


def test_reference_cli_links_and_reaches_our_library(tmp_path):
    """No GPU needed: the reference's main() parses its options, builds its net
    through our rnn_new / initialisers, and its first rnn_opinion lands in
    librecur_b200.so - which, without a CUDA device, refuses loudly."""
    need_cli()
    out = subprocess.run(["ldd", OUR_BIN], capture_output=True, text=True).stdout
    assert "librecur_b200.so" in out and "not found" not in out
    undefined = subprocess.run(["nm", "-D", "--undefined-only", OUR_BIN], capture_output=True,
                               text=True).stdout
    wanted = [l.split()[-1] for l in undefined.splitlines() if " rnn_" in l]
    assert {"rnn_opinion", "rnn_bptt_calculate", "rnn_new_with_bottom_layer",
            "rnn_new_training_set", "rnn_apply_learning"} <= set(wanted)
    exported = subprocess.run(["nm", "-D", "--defined-only",
                               os.path.join(ROOT, "recur_b200", "librecur_b200.so")],
                              capture_output=True, text=True).stdout
    have = {l.split()[-1] for l in exported.splitlines()}
    assert not [s for s in wanted if s not in have]
    from recur_b200 import api
    if api.load_library().rnn_b200_device_count() > 0:
        return      # the GPU test below does the real run
    r, _ = run_cli(OUR_BIN, str(tmp_path), 5)
    assert r.returncode != 0 and "no CPU compute path" in r.stderr


@pytest.mark.gpu
def test_reference_cli_config1_on_erewhon_matches_reference(gpu_lib, tmp_path):
    need_cli()
    n_gen = 4096
    rr, ref_log = run_cli(REF_BIN, str(tmp_path / "ref"), n_gen)
    gr, our_log = run_cli(OUR_BIN, str(tmp_path / "b200"), n_gen)
    assert rr.returncode == 0, rr.stderr[-2000:]
    assert gr.returncode == 0, gr.stderr[-2000:]
    want, want_rep = parse_log(ref_log)
    got, got_rep = parse_log(our_log)
    assert len(want) == len(got) == n_gen
    keys = ("depth", "ih_scale", "min_error_threshold", "min_error_factor", "cum_error",
            "hidden_sum", "hidden_magnitude", "hidden_zeros", "top_error_scaled",
            "top_error_raw", "error_sum", "scaled_error")
    assert set(keys) <= set(want[0]) and set(want[0]) == set(got[0])
    # step for step while rounding has not been amplified yet (five printed digits)
    for g in range(200):
        for key in keys:
            a, b = got[g][key], want[g][key]
            if key == "depth":
                assert a == b, (g, key, a, b)
            else:
                assert abs(a - b) <= 2e-3 * abs(b) + 1e-9, (g, key, a, b)
    # the whole run through its statistics, per 1024 generations (the two
    # trajectories part for good around generation 400; their means do not)
    for key, tol in (("depth", 0.02), ("hidden_zeros", 0.03), ("hidden_sum", 0.15),
                     ("top_error_raw", 0.15), ("cum_error", 0.2)):
        a = np.array([s[key] for s in got])
        b = np.array([s[key] for s in want])
        for lo in range(0, n_gen, 1024):
            ma, mb = a[lo:lo + 1024].mean(), b[lo:lo + 1024].mean()
            assert abs(ma - mb) < tol * abs(mb), (key, lo, ma, mb)
    # and the learning curve the program reports (charmodel-predict.c:350,376)
    assert len(got_rep) == len(want_rep) >= 3
    for a, b in zip(got_rep, want_rep):
        assert abs(a["t_entropy"] - b["t_entropy"]) < 0.01 * b["t_entropy"], (a, b)
        assert abs(a["accuracy"] - b["accuracy"]) < 0.02, (a, b)
    print("text-predict (the reference binary) chars/s: reference core %.0f, librecur_b200 %.0f"
          % (want_rep[-1]["per_second"], got_rep[-1]["per_second"]))
    assert want_rep[-1]["t_entropy"] < want_rep[0]["t_entropy"]   # it learns
