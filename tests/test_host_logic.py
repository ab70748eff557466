"""Host-side logic that must match the reference draw for draw / byte for
byte: random streams, weight initialisation, cloning, weight surgery, the
saved-net format.  None of this needs a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from recur_b200 import abi
from helpers import make_net, weights, arr, STD_FLAGS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURE = "/root/reference/test/multi-text-6c34c563i73-h99-o3650.net"


def rng_tuple(net):
    r = net.contents.rng
    return (r.a, r.b, r.c, r.d)


def test_rng_matches_golden_draws(lib):
    """recur-rng.h:24-43,62-75,179-200 against draws taken from the reference."""
    g = np.load(os.path.join(GOLDEN, "recur_golden.npz"))
    for k, seed in enumerate(g["rng_seeds"]):
        # a net's generator is seeded by rnn_new; draw through clones' subseeds
        net = lib.rnn_new(2, 2, 2, abi.RNN_NET_FLAG_OWN_WEIGHTS, int(seed), None, 0, 0, 0, 0, 1)
        st = net.contents.rng
        # replay the generator in Python from the seeded state
        def rot(x, k):
            return ((x << k) | (x >> (64 - k))) & (2**64 - 1)
        a, b, c, d = st.a, st.b, st.c, st.d
        outs = []
        for _ in range(8):
            e = (a - rot(b, 7)) % 2**64
            a = b ^ rot(c, 13)
            b = (c + rot(d, 37)) % 2**64
            c = (d + e) % 2**64
            d = (e + a) % 2**64
            outs.append(d)
        assert outs == [int(x) for x in g["rng_u64"][k]]
        lib.rnn_delete_net(net)


@pytest.mark.parametrize("shape", [abi.RNN_INIT_FLAT, abi.RNN_INIT_FAN_IN, abi.RNN_INIT_RUNS,
                                   abi.RNN_INIT_ZERO])
def test_init_methods_bit_exact(lib, ref, shape):
    a = make_net(lib, input_size=9, hidden=31, output=6, init=False)
    b = make_net(ref, input_size=9, hidden=31, output=6, init=False)
    lib.rnn_randomise_weights_simple(a, shape)
    ref.rnn_randomise_weights_simple(b, shape)
    for x, y in zip(weights(a), weights(b)):
        assert np.array_equal(x, y)
    assert rng_tuple(a) == rng_tuple(b)


@pytest.mark.parametrize("dist", [1, 2, 3, 4])
def test_flat_distributions_bit_exact(lib, ref, dist):
    nets = []
    for L in (lib, ref):
        n = make_net(L, input_size=5, hidden=23, output=4, init=False)
        p = abi.RecurInitialisationParameters()
        L.rnn_init_default_weight_parameters(n, C.byref(p))
        p.flat_shape = dist
        p.flat_perforation = 0.3
        L.rnn_randomise_weights_clever(n, C.byref(p))
        nets.append(n)
    for x, y in zip(weights(nets[0]), weights(nets[1])):
        if dist == 3:  # log-normal goes through fast_expf: last-bit FMA freedom
            np.testing.assert_allclose(x, y, rtol=2e-6, atol=0)
        else:
            assert np.array_equal(x, y)
    assert rng_tuple(nets[0]) == rng_tuple(nets[1])


def test_default_parameters_match(lib, ref):
    a = make_net(lib, input_size=42, hidden=199, output=42, init=False)
    b = make_net(ref, input_size=42, hidden=199, output=42, init=False)
    pa, pb = abi.RecurInitialisationParameters(), abi.RecurInitialisationParameters()
    lib.rnn_init_default_weight_parameters(a, C.byref(pa))
    ref.rnn_init_default_weight_parameters(b, C.byref(pb))
    for name, _ in abi.RecurInitialisationParameters._fields_:
        assert getattr(pa, name) == getattr(pb, name), name


def test_training_set_clones_share_and_subseed(lib, ref):
    out = []
    for L in (lib, ref):
        net = make_net(L)
        nets = L.rnn_new_training_set(net, 4)
        rngs = [rng_tuple(nets[j]) for j in range(4)]
        n0 = nets[0].contents
        for j in range(1, 4):
            c = nets[j].contents
            assert C.addressof(c.ih_weights.contents) == C.addressof(n0.ih_weights.contents)
            assert C.addressof(c.bptt.contents.ih_delta.contents) == C.addressof(n0.bptt.contents.ih_delta.contents)
            assert C.addressof(c.bptt.contents.ho_delta.contents) == C.addressof(n0.bptt.contents.ho_delta.contents)
            assert C.addressof(c.bptt.contents.ih_momentum.contents) == C.addressof(n0.bptt.contents.ih_momentum.contents)
            assert C.addressof(c.hidden_layer.contents) != C.addressof(n0.hidden_layer.contents)
            assert c.flags & abi.RNN_NET_FLAG_NO_DELTAS
            assert not (c.flags & abi.RNN_NET_FLAG_OWN_WEIGHTS)
            assert c.bptt.contents.index == 1
        out.append(rngs)
        L.rnn_delete_training_set(nets, 4, 0)
    assert out[0] == out[1]


def test_weight_surgery_bit_exact(lib, ref):
    res = []
    for L in (lib, ref):
        n = make_net(L, input_size=6, hidden=21, output=4, seed=9)
        L.rnn_perforate_weights(n, 0.5)
        L.rnn_perforate_weights(n, 0.25)
        L.rnn_weight_noise(n, 0.01)
        L.rnn_zap_non_diagonals(n, 10, 18, 2)
        L.rnn_clear_diagonal_only_section(n, 5, 2)
        L.rnn_clear_diagonal_only_section(n, 0, 0)  # the per-step no-op call
        res.append([w.copy() for w in weights(n)] + [rng_tuple(n)])
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    assert res[0][2] == res[1][2]


def test_scale_initial_weights_follows_reference_draws(lib, ref):
    res = []
    for L in (lib, ref):
        n = make_net(L, input_size=6, hidden=21, output=4, seed=5)
        L.rnn_scale_initial_weights(n, 0.7)
        res.append((weights(n)[0].copy(), rng_tuple(n)))
    assert res[0][1] == res[1][1]   # same number of draws, incl. the MAX() quirk
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-4, atol=1e-7)


def test_momentum_soft_start(lib, ref):
    for gen, m, x in ((0, 0.95, 2000.0), (1000, 0.95, 2000.0), (1e6, 0.95, 2000.0), (5, 0.5, 0.0)):
        assert lib.rnn_calculate_momentum_soft_start(gen, m, x) == \
            ref.rnn_calculate_momentum_soft_start(gen, m, x)


# ---- saved nets ----------------------------------------------------------------

def cdb_items(path):
    """Independent pure-Python CDB walk (format: cr.yp.to/cdb/cdb.txt)."""
    import struct
    data = open(path, "rb").read()
    tables = [struct.unpack_from("<LL", data, i * 8) for i in range(256)]
    end = min(p for p, n in tables)
    pos = 2048
    items = []
    while pos < end:
        klen, dlen = struct.unpack_from("<LL", data, pos)
        pos += 8
        items.append((data[pos:pos + klen], data[pos + klen:pos + klen + dlen]))
        pos += klen + dlen
    return items


def test_save_is_byte_identical_to_reference(lib, ref, tmp_path):
    """Same net, saved by both libraries: identical files (key order, raw
    values, hash tables)."""
    os.chdir(tmp_path)
    paths = []
    for name, L in (("ours.net", lib), ("ref.net", ref)):
        n = make_net(L, input_size=6, hidden=21, output=4, seed=5)
        n.contents.generation = 77
        assert L.rnn_save_net(n, name.encode(), 0) == 0
        paths.append(str(tmp_path / name))
    a, b = open(paths[0], "rb").read(), open(paths[1], "rb").read()
    assert cdb_items(paths[0]) == cdb_items(paths[1])
    assert a == b


def test_load_each_others_files(lib, ref, tmp_path):
    os.chdir(tmp_path)
    n = make_net(lib, input_size=6, hidden=21, output=4, seed=5)
    lib.rnn_save_net(n, b"ours.net", 0)
    m = ref.rnn_load_net(b"ours.net")
    assert m
    for x, y in zip(weights(n), weights(m)):
        assert np.array_equal(x, y)
    assert rng_tuple(n) == rng_tuple(m)
    g = os.path.join(GOLDEN, "ref_saved_small.net")
    k = lib.rnn_load_net(g.encode())
    r = ref.rnn_load_net(g.encode())
    assert k and r
    for x, y in zip(weights(k), weights(r)):
        assert np.array_equal(x, y)
    for f in ("generation", "flags", "activation", "input_size", "hidden_size", "output_size"):
        assert getattr(k.contents, f) == getattr(r.contents, f)
    for f in ("depth", "index", "learn_rate", "ho_scale", "momentum", "momentum_weight",
              "min_error_factor"):
        assert getattr(k.contents.bptt.contents, f) == getattr(r.contents.bptt.contents, f), f


def test_load_golden_net_without_reference(lib):
    """ref_saved_small.net was written by the reference's rnn_save_net
    (tests/golden/make_golden.py); its weights are the trace's final ones."""
    g = np.load(os.path.join(GOLDEN, "recur_golden.npz"))
    k = lib.rnn_load_net(os.path.join(GOLDEN, "ref_saved_small.net").encode())
    assert k
    ih, ho = weights(k)
    assert np.array_equal(ih, g["trace_ih_weights"][-1])
    assert np.array_equal(ho, g["trace_ho_weights"][-1])
    assert k.contents.generation == 40


def test_load_failure_returns_null(lib, tmp_path):
    assert not lib.rnn_load_net(str(tmp_path / "nope.net").encode())
    bad = tmp_path / "bad.net"
    bad.write_bytes(b"\0" * 4096)
    assert not lib.rnn_load_net(str(bad).encode())


@pytest.mark.skipif(not os.path.exists(FIXTURE), reason="reference fixture not on this machine")
def test_reference_fixture_loads_identically(lib, ref):
    """The one on-disk artefact the reference ships for this path
    (test/multi-text-6c34c563i73-h99-o3650.net, CDB format 10)."""
    a = lib.rnn_load_net(FIXTURE.encode())
    b = ref.rnn_load_net(FIXTURE.encode())
    assert a and b
    na, nb = a.contents, b.contents
    assert (na.input_size, na.hidden_size, na.output_size) == (73, 99, 3650)
    assert na.activation == abi.RNN_RESQRT and na.bptt.contents.depth == 50
    for x, y in zip(weights(a), weights(b)):
        assert np.array_equal(x, y)
    assert na.metadata == nb.metadata
    assert rng_tuple(a) == rng_tuple(b)
    keys = [k for k, v in cdb_items(FIXTURE)]
    assert keys[0] == b"save_format_version" and b"net.ih_weights" in keys


def test_stream_slots_are_recycled_and_sets_stay_contiguous(lib):
    """Host-side bookkeeping of the stream pools (rb_device.cu): slots freed
    by deleted clones are found again (next-fit hint), a training set made
    afterwards still works, and hundreds of clones taken one at a time (an
    rnnca grid makes one per pixel) do not disturb the nets made before."""
    net = make_net(lib, input_size=8, hidden=13, output=4, depth=5, seed=3)
    first = lib.rnn_new_training_set(net, 6)
    marks = []
    for j in range(6):
        c = first[j].contents
        arr(c.hidden_layer, c.h_size)[1] = 10.0 + j      # something to recognise the net by
        marks.append(C.addressof(c))
    # clones taken one by one: the pool grows several times underneath
    fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
    cells = [lib.rnn_clone(net, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(300)]
    assert all(cells)
    for j in range(6):
        c = first[j].contents
        assert arr(c.hidden_layer, c.h_size)[1] == 10.0 + j
        assert c.ih_weights and C.addressof(c) == marks[j]
    for cell in cells[::2]:
        lib.rnn_delete_net(cell)
    more = [lib.rnn_clone(net, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(200)]
    assert all(more)
    for cell in cells[1::2] + more:
        assert arr(cell.contents.hidden_layer, cell.contents.h_size)[1] == 0.0
        lib.rnn_delete_net(cell)
    second = lib.rnn_new_training_set(first[1], 4)       # clones of a clone share the same weights
    assert C.addressof(second[1].contents.ih_weights.contents) == \
        C.addressof(net.contents.ih_weights.contents)
    for j in range(1, 4):
        lib.rnn_delete_net(second[j])
    lib.rnn_delete_training_set(first, 6, 0)
