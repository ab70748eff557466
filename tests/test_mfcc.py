"""the audio front end of gstclassify (reference mfcc.c:9-94).

The reference's mfcc.c is compiled unmodified into oracle/_ref/libmfcc_ref.so
over a DFT standing in for GStreamer's FFT (oracle/shim/gst/fft/gstfftf32.h).
CPU half: that build against a numpy reading of the same pipeline (pins the
shim and the understanding), and this library's host-side set-up tables
against the reference's.  GPU half: rnn_mfcc_extract against the reference on
random and tonal windows, both output kinds, several window shapes."""
import ctypes as C
import os

import numpy as np
import pytest

from recur_b200 import abi

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                   "libmfcc_ref.so")

# gstclassify's defaults (gstclassify.c:90-106, 962-968) and variations
SHAPES = [
    dict(window=256, wtype=1, bins=32, fmin=60.0, fmax=8000 * 0.499, knee=700.0, focus=0.0,
         rate=8000.0, scale=1.0 / 32768, vsize=2),
    dict(window=512, wtype=2, bins=20, fmin=100.0, fmax=3500.0, knee=500.0, focus=600.0,
         rate=8000.0, scale=1.0 / 32768, vsize=2),
    dict(window=1024, wtype=3, bins=40, fmin=40.0, fmax=7000.0, knee=700.0, focus=0.0,
         rate=16000.0, scale=1.0, vsize=2),
    dict(window=128, wtype=0, bins=12, fmin=200.0, fmax=3900.0, knee=700.0, focus=0.0,
         rate=8000.0, scale=1.0 / 32768, vsize=2),
]


def _args(s):
    return (s["window"], s["wtype"], s["bins"], s["fmin"], s["fmax"], s["knee"], s["focus"],
            s["rate"], s["scale"], s["vsize"])


@pytest.fixture(scope="module")
def mref():
    if not os.path.exists(REF):
        if os.path.exists("/root/reference/mfcc.c"):
            import oracle
            oracle.build(ref=True, port=False)
        else:
            pytest.skip("oracle/_ref/libmfcc_ref.so not built and /root/reference absent")
    lib = C.CDLL(REF)
    lib.ref_mfcc_new.restype = C.c_void_p
    lib.ref_mfcc_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_float, C.c_int]
    lib.ref_mfcc_delete.argtypes = [C.c_void_p]
    lib.ref_mfcc_extract.argtypes = [C.c_void_p, fp, C.c_int, fp, C.c_int]
    lib.ref_mfcc_tables.argtypes = [C.c_void_p, fp, ip, ip, fp, fp, fp]
    return lib


def _tables(L, handle, window, bins, name):
    mask = np.zeros(window, dtype=np.float32)
    left = np.zeros(bins + 1, dtype=np.int32)
    right = np.zeros(bins + 1, dtype=np.int32)
    lf, rf, sl = (np.zeros(bins + 1, dtype=np.float32) for _ in range(3))
    getattr(L, name)(handle, mask.ctypes.data_as(fp), left.ctypes.data_as(ip),
                     right.ctypes.data_as(ip), lf.ctypes.data_as(fp), rf.ctypes.data_as(fp),
                     sl.ctypes.data_as(fp))
    return mask, left, right, lf, rf, sl


def _windows(s, n, seed):
    """PCM as gstclassify feeds it: s16 samples as floats (gstclassify.c:2025-2028)."""
    rs = np.random.RandomState(seed)
    N = s["window"]
    t = np.arange(N)
    rows = []
    for w in range(n):
        kind = w % 4
        if kind == 0:
            x = rs.randint(-20000, 20000, size=N)
        elif kind == 1:
            f = rs.uniform(0.01, 0.45)
            x = 12000 * np.sin(2 * np.pi * f * t + rs.uniform(0, 6)) + rs.randint(-50, 50, size=N)
        elif kind == 2:
            x = rs.randint(-3, 4, size=N)          # near silence: log(1 + tiny)
        else:
            x = 30000 * np.sign(np.sin(2 * np.pi * rs.uniform(0.02, 0.2) * t))
        rows.append(np.asarray(x, dtype=np.float32))
    amp = 1.0 if s["scale"] < 1 else 1.0 / 32768
    return np.ascontiguousarray(np.stack(rows) * amp, dtype=np.float32)


def _numpy_bins(s, tables, pcm):
    """recur_extract_log_freq_bins read again in numpy (float64 FFT)."""
    mask, left, right, lf, rf, sl = tables
    out = np.zeros((len(pcm), s["bins"]))
    for w, x in enumerate(pcm):
        xw = x.astype(np.float64) * (mask if s["wtype"] else 1.0)
        power = np.abs(np.fft.rfft(xw)) ** 2
        sum_left = 0.0
        for i in range(s["bins"] + 1):
            j = left[i]
            mul = sl[i] * lf[i]
            p = power[j] * lf[i]
            sum_right = sum_left + (1.0 - mul) * p
            sum_left = mul * p
            if left[i] != right[i]:
                for j in range(left[i] + 1, right[i]):
                    mul += sl[i]
                    sum_left += mul * power[j]
                    sum_right += (1.0 - mul) * power[j]
                j = right[i]
            mul += sl[i] * rf[i]
            p = power[j] * rf[i]
            sum_left += mul * p
            sum_right += (1.0 - mul) * p
            if i:
                out[w, i - 1] = np.log(sum_right + 1)
    return out


def _numpy_dct(bins):
    n = bins.shape[1]
    k = np.arange(n)
    out = np.stack([(bins * np.cos(np.pi / n * j * (k + 0.5))).sum(axis=1) for j in range(n)], axis=1)
    out[:, 0] *= 0.7071067811865476
    return out


@pytest.mark.parametrize("shape", SHAPES)
def test_reference_mfcc_build_against_numpy(mref, shape):
    """The unmodified mfcc.c over the DFT shim does what a float64 numpy
    reading of mfcc.c:9-94 and recur_dct (badmaths: the plain DCT-II that
    recur_dct_cached tabulates) does."""
    h = mref.ref_mfcc_new(*_args(shape))
    tables = _tables(mref, h, shape["window"], shape["bins"], "ref_mfcc_tables")
    pcm = _windows(shape, 12, 3)
    got = np.zeros((len(pcm), shape["bins"]), dtype=np.float32)
    mref.ref_mfcc_extract(h, pcm.ctypes.data_as(fp), len(pcm), got.ctypes.data_as(fp), 0)
    want = _numpy_bins(shape, tables, pcm)
    assert np.abs(got - want).max() < 2e-4 * max(1.0, np.abs(want).max())
    got_d = np.zeros_like(got)
    mref.ref_mfcc_extract(h, pcm.ctypes.data_as(fp), len(pcm), got_d.ctypes.data_as(fp), 1)
    want_d = _numpy_dct(want)
    assert np.abs(got_d - want_d).max() < 2e-4 * max(1.0, np.abs(want_d).max())
    mref.ref_mfcc_delete(h)


@pytest.mark.parametrize("shape", SHAPES)
def test_mfcc_setup_tables_match_reference(lib, mref, shape):
    """Window masks (mfcc.c:272-305) and slopes (:134-178, through the iterative
    mel -> Hz inverse) built by this library's host code against the
    reference's: the integer bin edges identical, the fractions to the last
    bits.  No device needed."""
    h = mref.ref_mfcc_new(*_args(shape))
    m = lib.rnn_mfcc_new(*_args(shape))
    assert m
    want = _tables(mref, h, shape["window"], shape["bins"], "ref_mfcc_tables")
    got = _tables(lib, m, shape["window"], shape["bins"], "rnn_mfcc_tables")
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    for g, w in zip((got[0],) + got[3:], (want[0],) + want[3:]):
        assert np.allclose(g, w, rtol=1e-6, atol=1e-9)
    lib.rnn_mfcc_delete(m)
    mref.ref_mfcc_delete(h)
    # sizes the kernel does not take are refused, not mangled
    bad = dict(shape, window=300)
    assert not lib.rnn_mfcc_new(*_args(bad))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dct", [0, 1])
def test_mfcc_on_device_matches_reference(gpu_lib, mref, shape, dct):
    """rnn_mfcc_extract for a batch of channels' windows against the reference
    run window by window.  Tolerance: the features are log(1 + power) of an
    FP32 FFT; 1e-4 of the row's scale (power sums are positive, there is no
    cancellation; the bins' DCT mixes signs, so it is scaled by the bins)."""
    lib = gpu_lib
    h = mref.ref_mfcc_new(*_args(shape))
    m = lib.rnn_mfcc_new(*_args(shape))
    n = 260     # more windows than gstclassify's 256 default channels; ragged vs warps
    pcm = _windows(shape, n, 11)
    want = np.zeros((n, shape["bins"]), dtype=np.float32)
    bins = np.zeros_like(want)
    mref.ref_mfcc_extract(h, pcm.ctypes.data_as(fp), n, want.ctypes.data_as(fp), dct)
    mref.ref_mfcc_extract(h, pcm.ctypes.data_as(fp), n, bins.ctypes.data_as(fp), 0)
    got = np.zeros_like(want)
    lib.rnn_mfcc_extract(m, pcm.ctypes.data_as(fp), n, got.ctypes.data_as(fp), dct)
    scale = np.maximum(1.0, np.abs(bins).max(axis=1, keepdims=True))
    assert (np.abs(got - want) / scale).max() < 1e-4
    assert np.abs(got).max() > 1.0      # not all zeros
    # a second call reuses the staging buffers; one window alone
    one = np.zeros(shape["bins"], dtype=np.float32)
    lib.rnn_mfcc_extract(m, pcm[5].ctypes.data_as(fp), 1, one.ctypes.data_as(fp), dct)
    assert np.array_equal(one, got[5])
    lib.rnn_mfcc_delete(m)
    mref.ref_mfcc_delete(h)
