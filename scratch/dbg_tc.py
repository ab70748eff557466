import sys, os, ctypes as C
ROOT='/root/repo'; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
import numpy as np
from recur_b200 import api, abi
from helpers import *
L = api.load_library()
shape = dict(input_size=42, hidden=int(os.environ.get('H','63')), output=42, depth=6)
n = int(os.environ.get('N','64'))
text = markov_text(3000, 42, seed=2)
def fwd_only(engine, steps=1):
    L.rnn_b200_set_engine(engine)
    net = make_net(L, seed=1, lr=2e-4, **shape)
    nets = L.rnn_new_training_set(net, n)
    b = L.rnn_batch_new(nets, n)
    for t in range(steps):
        L.rnn_batch_advance(b)
        hot = np.array([(t*7+j*3) % 42 for j in range(n)], dtype=np.uint8)
        L.rnn_batch_set_one_hot(b, u8ptr(hot))
        L.rnn_batch_opinion(b, 0.0)
    H1 = shape['hidden']+1
    hid = np.zeros((n, H1), dtype=np.float32)
    L.rnn_batch_get_hiddens(b, fptr(hid))
    # now one calc_deltas with a fixed error
    err = np.random.RandomState(0).randn(n, 42).astype(np.float32)*0.1
    L.rnn_batch_set_errors(b, fptr(err))
    L.rnn_batch_calc_deltas(b, 0)
    L.rnn_batch_pull(b)
    c = net.contents
    ihd = arr(c.bptt.contents.ih_delta, c.ih_size).copy().reshape(c.i_size, c.h_size)
    hod = arr(c.bptt.contents.ho_delta, c.ho_size).copy()
    herr = np.stack([arr(nets[j].contents.bptt.contents.h_error, c.i_size).copy() for j in range(n)])
    ierr = np.stack([arr(nets[j].contents.bptt.contents.i_error, c.i_size).copy() for j in range(n)])
    sc = np.array([nets[j].contents.bptt.contents.ih_scale for j in range(n)])
    L.rnn_batch_delete(b)
    return hid, ihd, hod, herr, ierr, sc
for steps in (1, 3):
    a = fwd_only(1, steps); b = fwd_only(2, steps)
    print('steps', steps, 'hidden rel', rel_err(b[0], a[0]), 'ih_delta rel', rel_err(b[1], a[1]), 'ho_delta', rel_err(b[2], a[2]),
          'h_error', rel_err(b[3], a[3]), 'i_error', rel_err(b[4], a[4]))
    d = np.abs(b[1]-a[1]); print('  ih_delta max abs', np.abs(a[1]).max(), 'worst row/col', np.unravel_index(d.argmax(), d.shape), 'rows with err', (d.max(1) > 1e-4*np.abs(a[1]).max()).sum(), 'cols', (d.max(0) > 1e-4*np.abs(a[1]).max()).sum())
    dh = np.abs(b[0]-a[0]); print('  hidden worst', np.unravel_index(dh.argmax(), dh.shape), 'bad cols', (dh.max(0) > 1e-5).sum(), 'bad rows', (dh.max(1)>1e-5).sum())
