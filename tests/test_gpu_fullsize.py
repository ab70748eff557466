"""BASELINE.json configs[1] at its full size (hidden 1023, 512 streams, depth
30), where the CPU oracle would need minutes per step: size-independent
properties instead of an element-wise oracle comparison.

* the two engines (tcgen05 3xTF32 and FP32 FMA) are independent
  implementations of the same contraction order-insensitive sums: they must
  agree to the parity tolerance;
* the tensor path is deterministic: two runs from the same state are bit-equal;
* linearity in the streams: when every stream reads the same symbols (periodic
  text whose period divides the stream spacing), 512 streams at learning rate
  lr/8 take the same step as 64 streams at lr - the summed delta is 8x."""
import ctypes as C

import numpy as np
import pytest

from recur_b200 import api
from helpers import make_net, weights, u8ptr, markov_text, rel_err
from test_gpu_tc import run_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4
SHAPE = dict(input_size=42, hidden=1023, output=42, depth=30)


def test_full_size_engines_agree_and_tensor_path_is_deterministic(gpu_lib):
    lib = gpu_lib
    text = markov_text(200000, 42, seed=2)
    steps, n, lr = 3, 512, 1e-6
    a = run_batch(lib, 2, SHAPE, n, steps, text, lr)
    b = run_batch(lib, 2, SHAPE, n, steps, text, lr)
    f = run_batch(lib, 1, SHAPE, n, steps, text, lr)
    for key in ("ih", "ho", "hidden", "ih_delta"):
        assert np.array_equal(a[key], b[key]), key          # bit for bit
    assert a["stats"] == b["stats"]
    # the update moved the weights: compare the movement, not the (dominant) initial values
    for key, w0 in (("ih", "ih0"), ("ho", "ho0")):
        da, df = a[key] - a[w0], f[key] - f[w0]
        assert np.abs(df).max() > 0
        assert rel_err(da, df) < 2e-3, key   # a difference of two nearly equal fp32 numbers
        assert rel_err(a[key], f[key]) < TOL, key
    assert rel_err(a["ih_delta"], f["ih_delta"]) < TOL
    assert rel_err(a["hidden"], f["hidden"]) < TOL
    assert a["stats"][2:] == f["stats"][2:]                  # correct, count
    assert abs(a["stats"][1] - f["stats"][1]) < TOL * abs(f["stats"][1])


def test_full_size_linearity_in_identical_streams(gpu_lib):
    lib = gpu_lib
    # every stream sees the same symbols: period 64, spacings (len-1)/512 and
    # (len-1)/64 are multiples of it
    period = markov_text(64, 42, seed=3)
    length = 64 * 512 * 4 + 1
    text = np.ascontiguousarray(np.resize(period, length))
    assert ((length - 1) // 512) % 64 == 0 and ((length - 1) // 64) % 64 == 0
    big = run_batch(lib, 2, SHAPE, 512, 2, text, 1e-5 / 8)
    small = run_batch(lib, 2, SHAPE, 64, 2, text, 1e-5)
    # identical streams inside each batch
    assert np.abs(big["hidden"] - big["hidden"][0]).max() == 0.0
    assert rel_err(big["ih_delta"], 8.0 * small["ih_delta"]) < TOL
    for key, w0 in (("ih", "ih0"), ("ho", "ho0")):
        assert rel_err(big[key], small[key]) < TOL, key
        assert rel_err(big[key] - big[w0], small[key] - small[w0]) < 2e-3, key
    assert big["stats"][3] == 8 * small["stats"][3]
    assert abs(big["stats"][1] - 8 * small["stats"][1]) < TOL * abs(big["stats"][1])
