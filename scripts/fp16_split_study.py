"""Numerical study for DESIGN.md's next step (CPU only, numpy): how close do
FP16 hi/lo operand planes (lo pre-scaled by 2^11 into a second accumulator)
come to the 3xTF32 planes this round uses, on data shaped like the weight
gradient's operands?  X: ring rows (bias 1, ReLU activations with half of them
zero, up to a few hundred), E: error rows whose magnitude falls over the BPTT
steps by five orders.  dW = X^T . E over 4096 rows, 192 x 160 outputs.
Reference: float64.  Errors are max |diff| / max |dW| (the parity metric)."""
import numpy as np

rs = np.random.RandomState(1)
R, I, H = 4096, 192, 160


def tf32(a):
    """round to nearest, ties away (cvt.rna.tf32.f32): keep 10 mantissa bits"""
    b = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def mm32(a, b):
    return (a.astype(np.float32).T @ b.astype(np.float32)).astype(np.float32)


X = np.maximum(rs.randn(R, I), 0) * rs.choice([0.2, 3.0, 150.0], size=(R, 1))
X[:, 0] = 1.0
decay = 10.0 ** (-5.0 * (np.arange(R) // (R // 16)) / 15.0)
E = rs.randn(R, H) * 0.3 * decay[:, None]
X = X.astype(np.float32)
E = E.astype(np.float32)
ref = X.astype(np.float64).T @ E.astype(np.float64)
scale = np.abs(ref).max()


def err(c):
    return np.abs(c.astype(np.float64) - ref).max() / scale


print("fp32 (numpy matmul):            %.2e" % err(mm32(X, E)))
xh, eh = tf32(X), tf32(E)
xl, el = tf32(X - xh), tf32(E - eh)
print("1xTF32:                         %.2e" % err(mm32(xh, eh)))
print("3xTF32 (this round):            %.2e" % err(mm32(xl, eh) + mm32(xh, el) + mm32(xh, eh)))
for s_e in (8.0, 1.0):
    xs, es = X * 1.0, E * np.float32(s_e)          # X needs no scale (soft clip bounds it)
    xh = xs.astype(np.float16)
    eh = es.astype(np.float16)
    xl = ((xs - xh.astype(np.float32)) * np.float32(2048)).astype(np.float16)
    el = ((es - eh.astype(np.float32)) * np.float32(2048)).astype(np.float16)
    main = mm32(xh, eh)
    cross = mm32(xl, eh) + mm32(xh, el)
    out = (main + cross / np.float32(2048)) / np.float32(s_e)
    print("3xFP16, E scale %g:            %.2e   (hi planes overflow: %s, zero hi entries of nonzero E: %.1f%%)"
          % (s_e, err(out), bool(np.isinf(xh).any() or np.isinf(eh).any()),
             100.0 * ((eh == 0) & (E != 0)).mean()))
xb = X.view(np.uint32) & np.uint32(0xFFFF0000)
xb = xb.view(np.float32)
print("bf16 truncation of X alone:     %.2e" % err(mm32(xb, E)))
