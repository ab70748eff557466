"""SURVEY.md §8 f4, the other half of rnnca's front end: recur_adaptive_downscale
(rescale.c:240-256), which remember_frame (gstrnnca.c:619-637) runs on every
plane of every incoming frame.  Byte work: the device result must equal the
reference's bit for bit, including the pixels its rounding leaves unwritten.
The oracle is the unmodified rescale.c compiled by oracle/Makefile
(oracle/_ref/librescale_ref.so)."""
import ctypes as C
import os

import numpy as np
import pytest

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                   "librescale_ref.so")
u8p = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module")
def rref():
    if not os.path.exists(REF):
        if os.path.exists("/root/reference/rescale.c"):
            import oracle
            oracle.build(ref=True, port=False)
        else:
            pytest.skip("oracle/_ref/librescale_ref.so not built and /root/reference absent")
    lib = C.CDLL(REF)
    lib.recur_adaptive_downscale.restype = None
    lib.recur_adaptive_downscale.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int,
                                             C.c_int]
    return lib


# (source w, h, stride pad) -> (dest w, h, stride pad)
CASES = [
    ((1920, 1080, 0), (144, 96, 0)),     # the element at 1080p: skipping mode, x13.3 / x11.25
    ((1920, 1080, 64), (144, 96, 16)),   # ... with padded strides
    ((960, 540, 0), (144, 96, 0)),       # chroma planes of 1080p 4:2:0: skipping
    ((640, 480, 0), (144, 96, 0)),       # skipping, non-integer factors
    ((576, 384, 0), (144, 96, 0)),       # exactly x4: the skipping threshold
    ((575, 383, 0), (144, 96, 0)),       # just under it: exact mode
    ((320, 240, 0), (144, 96, 0)),       # exact mode, x2.2 / x2.5
    ((200, 150, 8), (144, 96, 0)),       # exact mode, barely shrinking
    ((145, 97, 0), (144, 96, 0)),        # one row and one column to lose
    ((144, 96, 0), (144, 96, 0)),        # equal sizes: the plain copy
    ((1001, 777, 3), (77, 61, 5)),       # odd everything
    ((4096, 2160, 0), (1920, 1080, 0)),  # 4K to the 1080p automaton of BASELINE configs[4]
    ((3840, 2160, 0), (1920, 1080, 0)),  # exactly x2
]


@pytest.mark.gpu
@pytest.mark.parametrize("src_shape,dst_shape", CASES)
def test_adaptive_downscale_bit_exact(gpu_lib, rref, src_shape, dst_shape):
    lib = gpu_lib
    (sw, sh, spad), (dw, dh, dpad) = src_shape, dst_shape
    ss, ds = sw + spad, dw + dpad
    rs = np.random.RandomState(sw * 7 + dh)
    src = rs.randint(0, 256, size=ss * sh + 64).astype(np.uint8)
    # structure as well as noise: gradients and a bright block
    img = src[:ss * sh].reshape(sh, ss)
    img[:, :sw] = (img[:, :sw] // 4 + (np.arange(sw)[None, :] * 255 // sw) // 2 +
                   (np.arange(sh)[:, None] * 255 // sh) // 4).astype(np.uint8)
    img[sh // 3:sh // 2, sw // 4:sw // 2] = 255
    fill = rs.randint(0, 256, size=ds * dh + 64).astype(np.uint8)   # what is unwritten must survive
    want, got = fill.copy(), fill.copy()
    rref.recur_adaptive_downscale(src.ctypes.data_as(u8p), sw, sh, ss, want.ctypes.data_as(u8p),
                                  dw, dh, ds)
    r = lib.rnn_b200_adaptive_downscale(src.ctypes.data_as(C.c_void_p), sw, sh, ss,
                                        got.ctypes.data_as(C.c_void_p), dw, dh, ds)
    assert r == 0
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]
    assert not np.array_equal(got, fill)


@pytest.mark.gpu
def test_adaptive_downscale_refuses_what_the_reference_cannot(gpu_lib):
    lib = gpu_lib
    src = np.zeros(64 * 64, dtype=np.uint8)
    dst = np.zeros(128 * 128, dtype=np.uint8)
    # enlarging: the reference divides by zero samples
    assert lib.rnn_b200_adaptive_downscale(src.ctypes.data_as(C.c_void_p), 64, 64, 64,
                                           dst.ctypes.data_as(C.c_void_p), 128, 128, 128) == -1
    # more than 257 source rows per destination row: its 16-bit sums overflow
    tall = np.zeros(8 * 4000, dtype=np.uint8)
    assert lib.rnn_b200_adaptive_downscale(tall.ctypes.data_as(C.c_void_p), 8, 4000, 8,
                                           dst.ctypes.data_as(C.c_void_p), 2, 2, 2) == -1
