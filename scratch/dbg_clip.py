import sys, os, ctypes as C
ROOT='/root/repo'; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
import numpy as np
from recur_b200 import api, abi
from helpers import *
from test_gpu_tc import run_batch
L = api.load_library()
shape = dict(input_size=12, hidden=75, output=12, depth=8)
text = markov_text(2000, 12, seed=3)
for boost, lr, steps in ((2.5, 0.05, 12), (2.0, 0.02, 6), (2.5, 0.02, 5), (3.0, 0.01, 4), (2.0, 0.05, 8)):
    a = run_batch(L, 1, shape, 64, steps, text, lr, boost=boost)
    b = run_batch(L, 2, shape, 64, steps, text, lr, boost=boost)
    print(boost, lr, steps, 'clipped', (a['ih_scale']!=1).sum(), a['ih_scale'].min(), 'hidden max', a['hidden'].max(),
          'rel', [round(rel_err(b[k], a[k]),8) for k in ('hidden','ih','ho','ih_delta')])
