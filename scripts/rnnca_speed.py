"""config 5 at the reference's own size: 144 x 96 = 13,824 forward-only cells
(gstrnnca.h:13-51), one clone per cell sharing the trainers' weights; frames/s
of set_inputs -> opinion -> get_outputs through the batch API."""
import sys, time, ctypes as C
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from recur_b200 import api, abi
from helpers import make_net, fptr
lib = api.load_library()
W, Hh = 144, 96
if len(sys.argv) > 1:   # WxH, e.g. 1920x1080 (about a minute to make 2 M cells)
    W, Hh = (int(x) for x in sys.argv[1].split('x'))
n = W * Hh
a = make_net(lib, input_size=35, hidden=51, output=3, depth=10, seed=11, lr=3e-3)
fwd = abi.RNN_NET_FLAG_STANDARD & ~(abi.RNN_NET_FLAG_OWN_BPTT | abi.RNN_NET_FLAG_OWN_WEIGHTS)
t0 = time.perf_counter()
cells = [lib.rnn_clone(a, fwd, abi.RECUR_RNG_SUBSEED, None) for _ in range(n)]
t1 = time.perf_counter()
batch = lib.rnn_batch_new((abi.RecurNN_p * n)(*cells), n)
t2 = time.perf_counter()
print("%d clones in %.2f s, batch in %.2f s" % (n, t1 - t0, t2 - t1), flush=True)
big = n > 100000
rs = np.random.RandomState(5)
inputs = rs.random_sample((n, 35)).astype(np.float32)
outs = np.zeros((n, 3), dtype=np.float32)
frames = 200 if not big else 10
for f in range(5 if not big else 1):
    lib.rnn_batch_set_inputs(batch, fptr(inputs)); lib.rnn_batch_opinion(batch, 0.0); lib.rnn_batch_get_outputs(batch, fptr(outs))
t0 = time.perf_counter()
for f in range(frames):
    lib.rnn_batch_set_inputs(batch, fptr(inputs))
    lib.rnn_batch_opinion(batch, 0.0)
    lib.rnn_batch_get_outputs(batch, fptr(outs))
t1 = time.perf_counter()
print("%.1f frames/s (%.1f us per frame of %d cells, %.1f M cell-steps/s)" % (frames / (t1 - t0), (t1 - t0) / frames * 1e6, n, frames * n / (t1 - t0) / 1e6), flush=True)

# the same frame loop with the gather and the byte conversion on the device
# (rnn_batch_rnnca_frame): only the two 3-plane frames cross PCIe
off_y = np.array([(dx, dy) for dy in range(-2, 3) for dx in range(-2, 3)
                  if abs(dx) + abs(dy) <= 2 or (abs(dx), abs(dy)) == (2, 2)][:17], dtype=np.int32)
off_c = np.array([(dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)], dtype=np.int32)
frame = rs.randint(0, 256, size=3 * n).astype(np.uint8)
out = np.zeros(3 * n, dtype=np.uint8)
u8p = C.POINTER(C.c_uint8); ip = C.POINTER(C.c_int)
def frame_step():
    lib.rnn_batch_rnnca_frame(batch, frame.ctypes.data_as(u8p), out.ctypes.data_as(u8p), W, Hh,
                              off_y.ctypes.data_as(ip), 17, off_c.ctypes.data_as(ip), 8, 2, 0)
for f in range(5 if not big else 1): frame_step()
t0 = time.perf_counter()
for f in range(frames): frame_step()
t1 = time.perf_counter()
print("rnn_batch_rnnca_frame: %.1f frames/s (%.1f us per frame, %.1f M cell-steps/s)" % (frames / (t1 - t0), (t1 - t0) / frames * 1e6, frames * n / (t1 - t0) / 1e6))
