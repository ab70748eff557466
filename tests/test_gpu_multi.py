"""Two GPUs: streams sharded over two processes, deltas summed across them
(NCCL all-reduce and the fused peer-memory kernel), against the CPU oracle
replaying all streams.  Skipped on a one-GPU box."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_training_matches_oracle(gpu_lib):
    if gpu_lib.rnn_b200_device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("MULTIGPU_CHECK ")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(lines[-1][len("MULTIGPU_CHECK "):])
    for variant in ("nccl", "p2p"):
        assert out[variant]["replicas_equal"], variant
        assert out[variant]["ih_rel"] < 1e-4 and out[variant]["ho_rel"] < 1e-4, out
    # the cell automaton row-sharded over the ranks: bit-identical pictures
    assert out["cells_sharded_equal"]
    # a bottom layer in front: its deltas are summed across the ranks as well
    # (against the compiled reference replaying every rank's streams)
    if out.get("bottom") is not None:
        assert out["bottom"]["replicas_equal"]
        assert max(out["bottom"]["rel"]) < 1e-4, out["bottom"]
