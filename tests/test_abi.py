"""The drop-in boundary: symbol set and struct layouts."""
import ctypes as C
import os
import re

from recur_b200 import abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(rnn_[a-z0-9_]+)\s*\(", text))
    names -= {"rnn_log_float", "rnn_log_int"}  # static inline in the header
    return names


def test_library_exports_every_declared_symbol(lib):
    for header in ("recur-nn.h", "recur_b200.h"):
        for name in sorted(declared_functions(header)):
            assert hasattr(lib, name), "%s declared in %s but not exported" % (name, header)


def test_declared_symbol_lists_are_complete():
    assert declared_functions("recur-nn.h") == set(abi.RNN_API_SYMBOLS)
    assert declared_functions("recur_b200.h") - set(abi.RNN_API_SYMBOLS) == set(api.B200_API_SYMBOLS)


def test_reference_exports_the_same_rnn_symbols(ref):
    for name in abi.RNN_API_SYMBOLS:
        assert hasattr(ref, name), name


def test_struct_sizes_match_reference(ref):
    assert C.sizeof(abi.RecurNN) == ref.ref_sizeof_RecurNN()
    assert C.sizeof(abi.RecurNNBPTT) == ref.ref_sizeof_RecurNNBPTT()
    assert C.sizeof(abi.RecurExtraLayer) == ref.ref_sizeof_RecurExtraLayer()


def test_struct_sizes_are_the_documented_ones():
    # x86-64 SysV: recur-nn.h:158-227
    assert C.sizeof(abi.RecurNN) == 176
    assert C.sizeof(abi.RecurNNBPTT) == 128
    assert C.sizeof(abi.RecurExtraLayer) == 96


def test_new_net_fields_match_reference(lib, ref):
    from helpers import make_net
    for kw in (dict(), dict(input_size=42, hidden=199, output=42, depth=30),
               dict(input_size=3, hidden=4, output=1, depth=2)):
        a = make_net(lib, init=False, **kw).contents
        b = make_net(ref, init=False, **kw).contents
        for f in ("i_size", "h_size", "o_size", "input_size", "hidden_size",
                  "output_size", "ih_size", "ho_size", "flags", "generation",
                  "presynaptic_noise", "activation"):
            assert getattr(a, f) == getattr(b, f), f
        for f in ("a", "b", "c", "d"):
            assert getattr(a.rng, f) == getattr(b.rng, f)
        for f in ("depth", "index", "learn_rate", "ih_scale", "ho_scale",
                  "momentum", "momentum_weight", "min_error_factor"):
            assert getattr(a.bptt.contents, f) == getattr(b.bptt.contents, f), f
        # real_inputs sits hidden_size+1 floats into the current ring row
        off = C.addressof(a.real_inputs.contents) - C.addressof(a.input_layer.contents)
        assert off == 4 * (a.hidden_size + 1)
        hist = C.addressof(a.bptt.contents.history.contents)
        assert C.addressof(a.input_layer.contents) - hist == 4 * a.i_size * a.bptt.contents.index


def test_compute_without_device_fails_loudly(lib):
    """No CPU fallback: a compute call with no GPU must abort, not limp on."""
    import subprocess
    import sys
    if lib.rnn_b200_device_count() > 0:
        return
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from recur_b200 import api, abi\n"
            "L = api.load_library()\n"
            "n = L.rnn_new(3, 4, 2, abi.RNN_NET_FLAG_STANDARD, 1, None, 3, 0.1, 0.9, 0.0, 1)\n"
            "L.rnn_opinion(n, None, 0.0)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU compute path" in r.stderr
