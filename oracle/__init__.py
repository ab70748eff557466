"""TEST INFRASTRUCTURE ONLY.  Loaders for the two CPU oracles:

  ref   oracle/_ref/librecur_ref.so — the unmodified reference compiled in
        place (oracle/Makefile), driven through the same ctypes ABI as the
        product library;
  port  oracle/liboracle_rnn.so — this repo's plain-C restatement.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this package; the product (recur_b200/) never does.
"""
