"""config 1 (single net, H199, depth 30): chars/s through the per-net drop-in
API on the GPU vs the compiled reference on one host core.  Run from the repo
root on a GPU box: python scripts/config1_speed.py  (numbers: DESIGN.md)."""
import sys, time, ctypes as C
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import oracle
from recur_b200 import api, abi
from helpers import make_net, u8ptr, markov_text, arr
lib = api.load_library()
ref = oracle.load_ref(strict=False)
text = markov_text(20000, 42, seed=6)
shape = dict(input_size=42, hidden=199, output=42, depth=30, seed=1, lr=1e-3)
r = make_net(ref, **shape)
a = make_net(lib, **shape)
steps = 2000
t = ref.ref_single_net_train(r, u8ptr(text), len(text), 0, steps, 0.95, 2000.0, 1, None, None, None)
print("reference CPU: %.0f chars/s" % (steps / t))
c = a.contents
o_err = c.bptt.contents.o_error
def run(n0, n1):
    for i in range(n0, n1):
        c.bptt.contents.momentum = lib.rnn_calculate_momentum_soft_start(c.generation, 0.95, 2000.0)
        lib.rnn_bptt_advance(a)
        inputs = arr(c.real_inputs, c.input_size)
        inputs[:] = 0
        inputs[text[i]] = 1.0
        answer = lib.rnn_opinion(a, None, 0.0)
        ref.ref_softmax_best_guess(o_err, answer, c.output_size)
        o_err[text[i + 1]] += 1.0
        lib.rnn_bptt_calculate(a, 1)
run(0, 100)
t0 = time.perf_counter(); run(100, 600); t1 = time.perf_counter()
print("ours per-net API: %.0f chars/s (%.1f us/char), launches/char %.1f" % (500 / (t1 - t0), (t1 - t0) / 500 * 1e6, 0))
l0 = lib.rnn_b200_kernel_launches(); run(600, 700); print("launches per char", (lib.rnn_b200_kernel_launches() - l0) / 100)
import collections
tt = collections.defaultdict(float)
def run_timed(n0, n1):
    for i in range(n0, n1):
        t0 = time.perf_counter()
        c.bptt.contents.momentum = lib.rnn_calculate_momentum_soft_start(c.generation, 0.95, 2000.0)
        lib.rnn_bptt_advance(a)
        inputs = arr(c.real_inputs, c.input_size)
        inputs[:] = 0
        inputs[text[i]] = 1.0
        t1 = time.perf_counter()
        answer = lib.rnn_opinion(a, None, 0.0)
        t2 = time.perf_counter()
        ref.ref_softmax_best_guess(o_err, answer, c.output_size)
        o_err[text[i + 1]] += 1.0
        t3 = time.perf_counter()
        lib.rnn_bptt_calculate(a, 1)
        t4 = time.perf_counter()
        tt['prep'] += t1 - t0; tt['opinion'] += t2 - t1; tt['softmax'] += t3 - t2; tt['calculate'] += t4 - t3
run_timed(700, 1200)
print({k: round(v / 500 * 1e6, 1) for k, v in tt.items()})
lib.rnn_b200_profile_enable(1)
run(1200, 1700)
pms = (C.c_double * 8)(); pln = (C.c_uint64 * 8)()
ncls = lib.rnn_b200_profile_read(pms, pln, 8)
lib.rnn_b200_profile_enable(0)
for cidx in range(ncls):
    if pln[cidx]:
        print(lib.rnn_b200_profile_class_name(cidx).decode(), "%.1f us x %d" % (pms[cidx] / pln[cidx] * 1e3, pln[cidx]))
