import sys, os, tempfile
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import numpy as np
import test_reference_cli as t
d = tempfile.mkdtemp()
n = 4096
rr, rl = t.run_cli(t.REF_BIN, d + "/ref", n)
gr, gl = t.run_cli(t.OUR_BIN, d + "/b200", n)
w, wr = t.parse_log(rl); g, gr_ = t.parse_log(gl)
for key in ("depth", "hidden_zeros", "hidden_sum", "top_error_raw", "cum_error", "ih_scale"):
    a = np.array([s[key] for s in g]); b = np.array([s[key] for s in w])
    print(key, [(round(a[lo:lo+1024].mean(), 4), round(b[lo:lo+1024].mean(), 4)) for lo in range(0, n, 1024)])
first = next((i for i in range(n) if g[i]["depth"] != w[i]["depth"]), None)
print("first depth difference at generation", first)
for i in (10, 40, 100, 200, 400):
    print(i, {k: (g[i][k], w[i][k]) for k in ("hidden_sum", "cum_error", "top_error_raw")})
print(gr_); print(wr)
